"""Timing of one training step (Darknet-53 416x416, fp16x3 on tcgen05) at the per-GPU batch of BASELINE config 4 (16).
STEPS=n sets the timed steps (STEPS=1 WARM=1 under ncu; summarise the last `launches/step` rows with scripts/kernel_shares.py)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import yolo_b200  # noqa: E402
from yolo_b200 import synth  # noqa: E402

B = int(os.environ.get("B", "16"))
spec = {"size": [416, 416], "layers": [1, 2, 8, 8, 4], "channels": [32, 64, 128, 256, 512, 1024], "slice_point": [1, 3, 5, 6, 30],
        "all_anchors": [[[0.2216, 0.1552], [0.2144, 0.2408], [0.2825, 0.3456]], [[0.3959, 0.2706], [0.3703, 0.4351], [0.5708, 0.4278]],
                        [[0.4345, 0.6063], [0.5584, 0.7174], [0.7448, 0.6772]]], "classes": list(range(24)), "batch_size": B, "learning_rate": 0.001,
        "scale": {"score": 0.1, "box_yx": 0.01, "box_hw": 10.0, "rotate": 0.0, "class": 0.3}, "positive_weight": 1.0, "negative_weight": 0.1}
y = yolo_b200.YOLO(spec=spec, precision="fp16x3", max_batch=B)
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.uniform(0, 1, size=(B, 3, 416, 416)).astype(np.float32)).cuda()
synth.calibrated_params(y.net, x[:4].contiguous(), seed=1, channels_per_anchor=30)
lab = np.full((B, 1, 30), -1.0, np.float32)
lab[:, 0, :6] = [3, .5, .5, .3, .3, 0]
lab[:, 0, 6:] = 1.0 / 24
for _ in range(int(os.environ.get("WARM", "2"))):
    y._train_batch([x], [lab])
torch.cuda.synchronize()
t0 = time.perf_counter()
n = int(os.environ.get("STEPS", "5"))
for _ in range(n):
    y._train_batch([x], [lab])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
flops = 3 * 113.263e9 * B
print(f"train step B={B}: {dt*1e3:.1f} ms/step, {B/dt:.1f} img/s, ~{flops/dt/1e12:.1f} TFLOP/s (fwd+dgrad+wgrad, fp16x3 tcgen05), launches/step {y.net.launches}, "
      f"loss {float(y.last_losses.sum()):.4f}, saturation flags {y.net.saturated()}")
