"""Parity report on the BASELINE network (Darknet-53 416x416): every precision against the fp32 oracle and
against the oracle evaluated in float64 (run on the GPU box; writes a text table to stdout)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import yolo_b200  # noqa: E402
from gpu_util import oracle_outputs  # noqa: E402
from oracle import decode, nets, weights  # noqa: E402

B = int(os.environ.get("B", "2"))
spec = dict(nets.spec_dk53(), classes=list(range(24)))
params = weights.make_params("carnet", spec, seed=2024, calib_batch=1)
x, _ = weights.synthetic_frames(B, spec["size"], seed=1234)
ref32 = oracle_outputs("carnet", spec, params, x)
ref64 = oracle_outputs("carnet", spec, params, x, torch.float64)
p32, i32 = decode.predict(spec, ref32, return_index=True)
p64 = decode.predict(spec, [r.astype(np.float32) for r in ref64])
print(f"# Darknet-53 416x416, batch {B}, synthetic calibrated weights (seed 2024), frames seed 1234")
print(f"oracle fp32 vs its fp64 evaluation: heads max|d| = {max(np.abs(a - b).max() for a, b in zip(ref32, ref64)):.3e}; "
      f"predict rows[:, :5] max|d| = {np.abs(p32[:, :5] - p64[:, :5]).max():.3e}")
print("precision | heads max|cuda-f64| | heads max|cuda-f32oracle| | heads mean|cuda-f64| | rows[:5] max|cuda-f32oracle| | rows[:5] max|cuda-f64| | idx == oracle")
for prec in ("fp32", "fp16x3", "bf16x6", "bf16"):
    y = yolo_b200.YOLO(spec=spec, params=params, precision=prec, max_batch=B)
    out = [o.asnumpy() for o in y.net.forward(data=torch.from_numpy(x).cuda())]
    pred, idx = y.predict([torch.from_numpy(o).cuda() for o in out], return_index=True)
    e64 = max(np.abs(a - b).max() for a, b in zip(out, ref64))
    e32 = max(np.abs(a - b).max() for a, b in zip(out, ref32))
    m64 = np.mean([np.abs(a - b).mean() for a, b in zip(out, ref64)])
    print(f"{prec:7s} | {e64:.3e} | {e32:.3e} | {m64:.3e} | {np.abs(pred[:, :5] - p32[:, :5]).max():.3e} | "
          f"{np.abs(pred[:, :5] - p64[:, :5]).max():.3e} | {bool(np.array_equal(idx, i32))}")
    del y
