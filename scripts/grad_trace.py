"""Diagnostic (GPU box): activation gradients dy of every named layer after one training step against torch autograd (fp64 oracle),
in backward order - pinpoints the first layer whose incoming gradient is off.  Usage: python scripts/grad_trace.py [B] [H] [W]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import yolo_b200  # noqa: E402
from oracle import nets, train, weights  # noqa: E402
from test_gpu_train import spec_mid, _train_case  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
size = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (64, 64)
spec = spec_mid(size)
params, x, labels = _train_case(spec, B, 2, 31)
hp = train.V1_HPARAMS
taps = {}
tp = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=k.rsplit(".", 1)[1] in train.TRAINABLE) for k, v in params.items()}


def tap(name, t):
    t.retain_grad()
    taps[name] = t


out, _ = nets.forward("carnet", spec, tp, torch.as_tensor(x).double(), train=True, tap=tap)
targets, mask, _ = train.loss_mask(spec, labels)
losses = train.get_loss(spec, out, targets, mask, hp)
sum(l.sum() for l in losses).backward()
net = yolo_b200.Net("carnet", spec, precision="fp16x3", max_batch=B)
net.load_params(params)
tr = yolo_b200.Trainer(net)
tr.forward_backward(torch.from_numpy(x).cuda(), labels, hp["scale"], hp["positive_weight"], hp["negative_weight"])
print(f"B={B} size={size}: rel L2 error of dy per activation (backward order)")
for name in reversed(list(taps)):
    t = taps[name]
    if t.grad is None:
        continue
    try:
        got = net.activation("grad:" + name, tuple(t.shape))
    except yolo_b200.YoloError as e:
        print(f"  {name:30s} -- {str(e)[:60]}")
        continue
    g = t.grad.numpy()
    rel = np.linalg.norm(got - g) / max(np.linalg.norm(g), 1e-30)
    print(f"  {name:30s} {tuple(t.shape)!s:20s} rel {rel:.2e}  |dy| {np.linalg.norm(g):.2e}")
