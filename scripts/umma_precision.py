"""Error statistics of the tensor-core paths on single layers (run on the GPU box).  Prints signed-error
statistics against float64 so that accumulator truncation (round-toward-zero bias growing with the number
of MMA accumulation steps) can be told apart from random rounding noise."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from test_gpu_conv import run_layer  # noqa: E402

for cin, k in [(64, 1), (256, 1), (1024, 1), (64, 3), (256, 3), (1024, 3)]:
    for prec in ("fp32", "fp16x3", "bf16x6", "bf16"):
        out, ref, _ = run_layer(prec, 2, 16, 16, cin, 128, k, 1, k // 2, act=0, bn=0, seed=1)
        e = (out - ref)
        rel_bias = float(np.mean(e * np.sign(ref)) / np.mean(np.abs(ref)))     # < 0: magnitudes shrink (truncation)
        print(f"K={cin*k*k:5d} {prec:7s} max|e|={np.abs(e).max():.3e} rms={np.sqrt(np.mean(e**2)):.3e} "
              f"signed-rel-bias={rel_bias:+.3e} mean|ref|={np.mean(np.abs(ref)):.3f}")
