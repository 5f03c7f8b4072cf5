#!/usr/bin/env python
"""Benchmark of the YOLOv3 detection hot path (BASELINE.json: 416x416 images/s).

A "step" = one pass of the hot path over one batch of synthetic frames: Darknet-53 416x416 forward
(75 fused conv launches) + fused decode/top-1 (1 launch), batch 32 per GPU (BASELINE config[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16x6|fp32|bf16] [--batch B]
  python bench.py --impl reference ...      # the restated CPU path of the reference (oracle/), host cores

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same metric
through the reference-facing Python calls (net.forward + predict -> numpy) from pinned HOST uint8 frames.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "416x416 images/sec (Darknet-53 forward + fused decode/top-1)"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=None)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step probe")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload_config(args):
    return {"workload": f"Darknet-53 (layers [1,2,8,8,4], channels [32..1024], C=30) {args.size}x{args.size} "
                        f"inference, batch {args.batch}/GPU, decode+top-1 fused (BASELINE configs[1])",
            "batch_per_gpu": args.batch, "size": [args.size, args.size]}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d, "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self, wait_s=3.0):
        """Call right before the timed region: waits until nvidia-smi delivers (its start-up takes driver locks that
        can stall a kernel submission for tens of ms - keep that out of the timed steps) and drops the earlier samples."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < wait_s:
            time.sleep(0.02)
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()            # the exact process we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in (self.lines[self.first:] or self.lines[-1:]):      # very short runs: fall back to the latest sample
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(args, steps, warmup, seconds_budget):
    """The reference's own CPU path, restated (oracle/): forward + predict with all host threads, timed the
    reference's way (yolo_modules/yolo_gluon.py:317-331: warm-ups, then N timed forwards each synchronised)."""
    import numpy as np
    import torch
    from oracle import decode, nets, weights
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = nets.spec_dk53((args.size, args.size))
    params = weights.to_torch(weights.make_params("carnet", spec, seed=2024, calib_batch=1))
    sample_b = 2
    x = torch.from_numpy(weights.synthetic_frames(sample_b, spec["size"], seed=1234)[0])
    def step():
        with torch.no_grad():
            heads = nets.forward("carnet", spec, params, x)
        return decode.predict(spec, [h.numpy() for h in heads])
    for _ in range(max(1, warmup)):
        step()
    t0 = time.perf_counter()
    n = 0
    while n < steps and (n == 0 or time.perf_counter() - t0 < seconds_budget):
        step()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": sample_b * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} steps x {sample_b} frames of the same {args.size}x{args.size} Darknet-53 workload (torch-CPU fp32/oneDNN "
                      f"restatement of the MXNet path, not MXNet), {dt:.1f} s"}, dt / max(n, 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, ms = cpu_reference_run(args, args.steps, args.warmup, 150.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def train_probe(spec, local, batch=16, steps=3):
    """Secondary measurement (BASELINE configs[3] per-GPU batch): one data-parallel training step = train-mode forward, targets +
    losses, backward, Adam (fp32 FFMA kernels), timed with CUDA events.  Never fails the bench: errors are reported as text."""
    try:
        import numpy as np
        import torch

        import yolo_b200
        from yolo_b200 import synth
        tspec = dict(spec, batch_size=batch, learning_rate=0.001, scale={"score": 0.1, "box_yx": 0.01, "box_hw": 10.0, "rotate": 0.0, "class": 0.3},
                     positive_weight=1.0, negative_weight=0.1)
        y = yolo_b200.YOLO(spec=tspec, precision="fp32", max_batch=batch, gpu=local)
        y.net.load_params(synth.random_params(y.net.param_shapes(), seed=1, channels_per_anchor=30))
        S = spec["size"][0]
        x = torch.rand((batch, 3, S, S), device=f"cuda:{local}")
        lab = np.full((batch, 1, 30), -1.0, np.float32)
        lab[:, 0, :6] = [3, .5, .5, .3, .3, 0]
        lab[:, 0, 6:] = 1.0 / 24
        y._train_batch([x], [lab])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            y._train_batch([x], [lab])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {"images_per_s": batch / (ms / 1e3), "ms_per_step": ms, "batch_per_gpu": batch, "dtype": "f32 (FFMA kernels)",
               "tflops": 3 * y.net.conv_flops_per_image * batch / (ms / 1e3) / 1e12, "launches_per_step": y.net.launches + 1,
               "what": "train-mode forward + targets/losses + backward + Adam (car/YOLO.py:350-399), no all-reduce at N=1"}
        del y
        torch.cuda.empty_cache()
        return out
    except Exception as e:          # noqa: BLE001
        return {"error": repr(e)[:300]}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import yolo_b200
    from yolo_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    precision = args.precision or "fp16x3"
    B, S = args.batch, args.size
    spec = {"size": [S, S], "layers": [1, 2, 8, 8, 4], "channels": [32, 64, 128, 256, 512, 1024], "slice_point": [1, 3, 5, 6, 30],
            "all_anchors": [[[0.2216, 0.1552], [0.2144, 0.2408], [0.2825, 0.3456]], [[0.3959, 0.2706], [0.3703, 0.4351], [0.5708, 0.4278]],
                            [[0.4345, 0.6063], [0.5584, 0.7174], [0.7448, 0.6772]]],
            "classes": list(range(24)), "use_fp16": False}
    y = yolo_b200.YOLO(args=None, spec=spec, precision=precision, max_batch=B, gpu=local)
    # every rank = an independent replica with the same weights (inference shards by batch, no collective)
    y.net.load_params(synth.random_params(y.net.param_shapes(), seed=2024, channels_per_anchor=30))
    rng = np.random.default_rng(1234 + rank)
    frames_u8 = torch.from_numpy(rng.integers(0, 256, size=(B, S, S, 3), dtype=np.uint8))
    host_frames = frames_u8.pin_memory()
    x_dev = (frames_u8.to(dev).permute(0, 3, 1, 2).float() / 255.0).contiguous()      # cv_img_2_ndarray layout, resident in HBM

    stream = torch.cuda.current_stream(dev)
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        out = y.net.forward(is_train=False, data=x_dev)
        rows, idx = yolo_b200.decode_top1(spec, out, y.steps)
        return out, rows, idx

    # ---- value: inputs resident in HBM ------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step_resident()
    launches_per_step = y.net.launches + 1
    barrier()
    if rank == 0:
        sampler.mark()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 1)]
    barrier()
    t_wall0 = time.perf_counter()
    ev[0].record(stream)
    for i in range(args.steps):
        out = y.net.forward(is_train=False, data=x_dev)
        ev[3 * i + 1].record(stream)
        rows, idx = yolo_b200.decode_top1(spec, out, y.steps)
        ev[3 * i + 2].record(stream)
        ev[3 * i + 3].record(stream)
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[3 * args.steps])
    fwd_ms = sum(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(args.steps)) / args.steps
    dec_ms = sum(ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(args.steps)) / args.steps
    t = torch.tensor([total_ms, fwd_ms, dec_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, fwd_ms, dec_ms = t.tolist()
    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms / 1e3)

    # ---- e2e: host uint8 frames -> H2D -> forward -> predict -> numpy rows, every step --------------------------
    for _ in range(2):
        y.predict(y.net.forward(is_train=False, data=host_frames))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        pred = y.predict(y.net.forward(is_train=False, data=host_frames))
    e1.record(stream)
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (e2e_ms.item() / 1e3)
    assert pred.shape == (B, 30)

    if rank == 0:
        pk, pk_src = peaks()
        flops_img = y.net.conv_flops_per_image
        passes = {"fp16x3": 3, "bf16x6": 6, "bf16": 1, "fp32": 1}[precision]
        if precision == "fp32":
            peak_tf, peak_note = 72.0, "fp32 FFMA nominal (148 SM x 128 lanes x 2 x ~1.9 GHz); not a tensor-core kernel"
        else:
            peak_tf = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1400.0))) / passes
            peak_note = (f"{pk_src} bf16 sustained {pk.get('bf16_tflops_sustained', pk.get('bf16_tflops'))} TF/s / {passes} "
                         f"tcgen05 16-bit MMA passes per fp32-grade product" if passes > 1 else f"{pk_src} bf16 sustained")
        achieved = B * flops_img / (fwd_ms / 1e3) / 1e12
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, "profiles", f"r1_step_traffic_{precision}.json")
        if os.path.exists(tp) and B == 32 and S == 416:
            tj = json.load(open(tp))
            traffic = tj["conv_launches_dram_bytes_per_step"]
            traffic_note = ("dram__bytes_read+write summed over the conv launches of one step from a committed ncu capture "
                            "(profiles/r1_step_traffic_%s.csv), not measured in this run" % precision)
        dec_bytes = B * sum(o.t[0].numel() for o in out) * 4
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16x3": "f32 (emulated: 3 fp16 tcgen05 passes on 2-way split operands, fp32 accumulate)",
                      "bf16x6": "f32 (emulated: 6 bf16 tcgen05 passes on 3-way split operands, fp32 accumulate)", "fp32": "f32",
                      "bf16": "bf16"}[precision],
            "data": "synthetic", "config": dict(workload_config(args), precision=precision, parallelism=f"replicas x{world} (no collective)",
                                                l2="per-step activation working set (GBs) exceeds the 126 MB L2; no explicit flush"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host_frames.numel()), "d2h_bytes_per_step": int(B * 30 * 4),
                    "api": "YOLO.net.forward(data=pinned uint8 NHWC host frames) + YOLO.predict -> numpy"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": traffic, "traffic_note": traffic_note,
                         "kernel": "conv forward (all conv launches of a step)", "algorithmic_flops_per_step": B * flops_img,
                         "avg_forward_ms": fwd_ms, "peak_source": peak_note},
            "roofline_decode": {"bound": "hbm", "achieved": dec_bytes / (dec_ms / 1e3) / 1e9, "peak": float(pk.get("hbm_gbs", 6650.0)),
                                "unit": "GB/s", "frac": dec_bytes / (dec_ms / 1e3) / 1e9 / float(pk.get("hbm_gbs", 6650.0)),
                                "algorithmic_bytes_per_step": dec_bytes, "avg_decode_ms": dec_ms, "peak_source": pk_src,
                                "note": "heads were just written by the head convs (L2-resident); launch-latency bound at this size"},
            "clocks": clocks, "wall_s": wall,
            "step_ms_each": [round(ev[3 * i].elapsed_time(ev[3 * i + 3]), 3) for i in range(args.steps)],
        }
        if world == 1 and not args.no_train:
            line["train_step"] = train_probe(spec, local)
        if world == 1 and not args.no_cpu_baseline:
            del y
            torch.cuda.empty_cache()
            line["cpu_baseline"], _ = cpu_reference_run(args, 1000, 1, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
