"""Fused decode kernels on COLD head tensors (5 rotating 40.9 MB sets > the 126 MB L2), for ncu captures:
    ncu --set full --clock-control none -k regex:decode_kernel -s 20 -c 3 -o gpurun_out/r2_decode python scripts/decode_probe.py
Launch order after the 20 warm-up launches: top-1, NMS at ~100 candidates / image, NMS at ~1000 candidates / image, repeating."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import yolo_b200  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from bench import dk53_spec  # noqa: E402

spec = dk53_spec(416)
B, nrot = 32, 5
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
sets = []
for _ in range(nrot):
    hs = []
    for hw in (52 * 52, 26 * 26, 13 * 13):
        t = torch.randn((B, hw, 3, 30), device=dev, generator=g)
        t[..., 0] = t[..., 0] * 2 - 4
        t[..., 3:5] *= 0.5
        hs.append(t)
    sets.append(hs)
sc = torch.sigmoid(torch.cat([t[..., 0].reshape(B, -1) for t in sets[0]], dim=1))
thr = {n: float(torch.topk(sc, n, dim=1).values[:, -1].mean()) for n in (100, 1000)}
k = 0
for _ in range(20):
    yolo_b200.decode_top1(spec, sets[k % nrot]); k += 1
for _ in range(int(os.environ.get("REPS", "4"))):
    yolo_b200.decode_top1(spec, sets[k % nrot]); k += 1
    yolo_b200.decode_nms(spec, sets[k % nrot], thr[100], 0.45, 100, 1024); k += 1
    yolo_b200.decode_nms(spec, sets[k % nrot], thr[1000], 0.45, 100, 1024); k += 1
torch.cuda.synchronize()
print("thresholds", thr)
