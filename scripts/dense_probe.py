"""LPDenseNet v2 (licence_plate/v2/spec.yaml: 320x512, growth 16, blocks [6,12,24,16]) at batch B: a few forwards for
ncu --metrics gpu__time_duration.sum (summarise with scripts/kernel_shares.py) and an event timing without the profiler."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import yolo_b200  # noqa: E402
from yolo_b200 import synth  # noqa: E402

B = int(os.environ.get("B", "64"))
spec = {"size": [320, 512], "num_init_features": 64, "growth_rate": 16, "block_config": [6, 12, 24, 16], "bn_size": 4,
        "LP_slice_point": [1, 3, 4, 7, 10], "LP_r_max": [45.0, 60.0, 45.0], "LP_num_class": 3}
lp = yolo_b200.LicencePlateDetectioin(spec=spec, precision="fp16x3", max_batch=B, gpu=0)
lp.net.load_params(synth.random_params(lp.net.param_shapes(), seed=5))
x = torch.rand((B, 3, 320, 512), device="cuda")
for _ in range(int(os.environ.get("WARM", "2"))):
    lp.net.forward(data=x)
torch.cuda.synchronize()
n = int(os.environ.get("STEPS", "5"))
t0 = time.perf_counter()
for _ in range(n):
    lp.net.forward(data=x)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"LPDenseNet B={B}: {dt*1e3:.2f} ms/forward, {B/dt:.0f} img/s, launches {lp.net.launches}, conv GFLOP/img {lp.net.conv_flops_per_image/1e9:.2f}")
