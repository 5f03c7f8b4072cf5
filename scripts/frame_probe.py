"""Single 416x416 frame (BASELINE configs[0]) through Darknet-53 + decode/top-1: a few forwards for an ncu launch list
(scripts/kernel_shares.py) and an event timing without the profiler."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import yolo_b200  # noqa: E402
from yolo_b200 import synth  # noqa: E402

spec = {"size": [416, 416], "layers": [1, 2, 8, 8, 4], "channels": [32, 64, 128, 256, 512, 1024], "slice_point": [1, 3, 5, 6, 30],
        "all_anchors": [[[0.2216, 0.1552], [0.2144, 0.2408], [0.2825, 0.3456]], [[0.3959, 0.2706], [0.3703, 0.4351], [0.5708, 0.4278]],
                        [[0.4345, 0.6063], [0.5584, 0.7174], [0.7448, 0.6772]]], "classes": list(range(24))}
y = yolo_b200.YOLO(spec=spec, precision="fp16x3", max_batch=1)
y.net.load_params(synth.random_params(y.net.param_shapes(), seed=2024, channels_per_anchor=30))
x = torch.from_numpy(np.random.default_rng(1).integers(0, 256, size=(1, 416, 416, 3), dtype=np.uint8)).cuda()
for _ in range(int(os.environ.get("WARM", "3"))):
    yolo_b200.decode_top1(spec, y.net.forward(data=x), y.steps)
torch.cuda.synchronize()
n = int(os.environ.get("STEPS", "20"))
t0 = time.perf_counter()
for _ in range(n):
    yolo_b200.decode_top1(spec, y.net.forward(data=x), y.steps)
torch.cuda.synchronize()
print(f"single frame: {(time.perf_counter() - t0) / n * 1e3:.3f} ms, launches {y.net.launches + 1}")
