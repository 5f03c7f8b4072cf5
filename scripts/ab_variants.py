"""In-process A/B timing of kernel variants selected by environment switches (read at every launch by the library),
so that all variants see the same memory placement, box and clocks.  Run on the GPU box:
    python scripts/ab_variants.py "YOLO_B200_WIDE=0" "YOLO_B200_WIDE=1" ...   (each argument: comma-separated KEY=VAL list)
Prints the median step time (Darknet-53 416, batch 32, forward only) of every variant over interleaved rounds."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import yolo_b200  # noqa: E402
from yolo_b200 import synth  # noqa: E402

variants = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[1:]] or [{}]
keys = sorted({k for v in variants for k in v})
B = int(os.environ.get("B", "32"))
spec = {"size": [416, 416], "layers": [1, 2, 8, 8, 4], "channels": [32, 64, 128, 256, 512, 1024], "slice_point": [1, 3, 5, 6, 30],
        "all_anchors": [[[0.2216, 0.1552], [0.2144, 0.2408], [0.2825, 0.3456]], [[0.3959, 0.2706], [0.3703, 0.4351], [0.5708, 0.4278]],
                        [[0.4345, 0.6063], [0.5584, 0.7174], [0.7448, 0.6772]]],
        "classes": list(range(24)), "use_fp16": False}
# the pseudo-key LIB=<path of another build of the library> gives that variant its own network from that build
from yolo_b200 import _lib as libmod  # noqa: E402
nets_by_lib = {}
for v in variants:
    path = v.pop("LIB", None)
    if path not in nets_by_lib:
        if path:
            libmod._lib, libmod.LIB_PATH = None, os.path.abspath(path)
        n = yolo_b200.YOLO(args=None, spec=spec, precision=os.environ.get("PREC", "fp16x3"), max_batch=B).net
        n.load_params(synth.random_params(n.param_shapes(), seed=2024, channels_per_anchor=30))
        nets_by_lib[path] = n
    v["_net"] = nets_by_lib[path]
keys = sorted({k for v in variants for k in v if k != "_net"})
x = torch.rand(B, 3, 416, 416, device="cuda")
rounds, steps = int(os.environ.get("ROUNDS", "4")), int(os.environ.get("STEPS", "5"))
times = [[] for _ in variants]
for r in range(rounds + 1):
    for i, v in enumerate(variants):
        for k in keys:
            os.environ.pop(k, None)
        net = v["_net"]
        os.environ.update({k: x_ for k, x_ in v.items() if k != "_net"})
        net.forward(is_train=False, data=x)
        torch.cuda.synchronize()
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            net.forward(is_train=False, data=x)
            e1.record()
            torch.cuda.synchronize()
            if r > 0:
                times[i].append(e0.elapsed_time(e1))
for a, t in zip(sys.argv[1:] or [""], times):
    print(f"{a:60s} median {np.median(t):7.3f} ms  min {np.min(t):7.3f}  max {np.max(t):7.3f}")
