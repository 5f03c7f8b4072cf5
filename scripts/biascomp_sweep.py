"""Truncation-bias compensation sweep (run on the GPU box): Darknet-53 416x416 head error against the fp64 oracle for several values
of YOLO_B200_BIASCOMP (relative gain delta * 2^-23 per 8 full-magnitude MMA additions applied when a TMEM partial is drained, see conv_umma.cu) and, optionally, YOLO_B200_HHLAST.
The oracle is evaluated once; every setting runs in a fresh process (the switches are read once per process)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
CACHE = "/tmp/biascomp_oracle.npz"

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import yolo_b200
    from oracle import nets, weights
    d = np.load(CACHE)
    spec = dict(nets.spec_dk53(), classes=list(range(24)))
    params = weights.make_params("carnet", spec, seed=2024, calib_batch=1)
    y = yolo_b200.YOLO(spec=spec, params=params, precision="fp16x3", max_batch=2)
    out = [o.asnumpy() for o in y.net.forward(data=torch.from_numpy(d["x"]).cuda())]
    ref64 = [d[f"r64_{i}"] for i in range(3)]
    ref32 = [d[f"r32_{i}"] for i in range(3)]
    e64 = max(np.abs(a - b.reshape(a.shape)).max() for a, b in zip(out, ref64))
    e32 = max(np.abs(a - b.reshape(a.shape)).max() for a, b in zip(out, ref32))
    rms = np.sqrt(np.mean(np.concatenate([(a - b.reshape(a.shape)).ravel() ** 2 for a, b in zip(out, ref64)])))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    xd = torch.from_numpy(d["x"]).cuda()
    for _ in range(3):
        y.net.forward(data=xd)
    ev[0].record()
    for _ in range(10):
        y.net.forward(data=xd)
    ev[1].record()
    torch.cuda.synchronize()
    print(f"BIASCOMP={os.environ.get('YOLO_B200_BIASCOMP', '-'):>5s} HHLAST={os.environ.get('YOLO_B200_HHLAST', '-')} "
          f"heads max|cuda-f64|={e64:.3e} max|cuda-f32oracle|={e32:.3e} rms|cuda-f64|={rms:.3e} fwd(B=2)={ev[0].elapsed_time(ev[1]) / 10:.3f} ms", flush=True)
    sys.exit(0)

import torch  # noqa: E402
from gpu_util import oracle_outputs  # noqa: E402
from oracle import nets, weights  # noqa: E402

spec = dict(nets.spec_dk53(), classes=list(range(24)))
params = weights.make_params("carnet", spec, seed=2024, calib_batch=1)
x, _ = weights.synthetic_frames(2, spec["size"], seed=1234)
r32 = oracle_outputs("carnet", spec, params, x)
r64 = oracle_outputs("carnet", spec, params, x, torch.float64)
np.savez(CACHE, x=x, **{f"r32_{i}": r for i, r in enumerate(r32)}, **{f"r64_{i}": r for i, r in enumerate(r64)})
print(f"oracle fp32 vs fp64: heads max|d| = {max(np.abs(a - b).max() for a, b in zip(r32, r64)):.3e}", flush=True)
settings = [a for a in sys.argv[1:]] or ["0", "1.0", "1.4", "1.8", "2.2"]
for s in settings:
    env = dict(os.environ)
    parts = s.split(":")
    env["YOLO_B200_BIASCOMP"] = parts[0]
    if len(parts) > 1:
        env["YOLO_B200_HHLAST"] = parts[1]
    subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=False)
