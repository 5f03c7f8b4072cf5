"""Multi-tile-per-CTA regression check for the persistent tcgen05 kernel (more tiles than SMs so that both
epilogue groups and the partial ring wrap several times).  Run on the GPU box under `timeout`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from test_gpu_conv import run_layer  # noqa: E402

for prec in ("fp16x3", "bf16x6", "bf16"):
    for (B, H, W, cin, cout, k, s, p, res) in [(8, 52, 52, 64, 128, 3, 1, 1, 0), (4, 104, 104, 64, 64, 3, 1, 1, 1), (4, 104, 104, 32, 64, 3, 1, 1, 0), (16, 26, 26, 256, 256, 1, 1, 0, 0),
                                               (6, 40, 40, 128, 90, 1, 1, 0, 0), (4, 26, 26, 256, 512, 3, 1, 1, 0), (8, 13, 13, 512, 512, 3, 1, 1, 1),
                                               (32, 26, 26, 256, 512, 3, 1, 1, 0), (16, 52, 52, 256, 256, 3, 1, 1, 1), (32, 26, 26, 512, 256, 1, 1, 0, 0)]:
        out, ref, _ = run_layer(prec, B, H, W, cin, cout, k, s, p, act=1, residual=res, bn=1, seed=3)
        e = np.abs(out - ref).max()
        tiles = ((B * out.shape[1] * out.shape[2] + 127) // 128) * ((cout + 127) // 128)
        print(f"{prec:7s} tiles={tiles:5d} cin={cin:4d} cout={cout:4d} k={k} max|e|={e:.3e}", flush=True)
        assert e < (0.1 if prec == "bf16" else 2e-5), "mismatch"
print("multitile OK")
