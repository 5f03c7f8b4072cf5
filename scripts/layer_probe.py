"""Single-layer probe for the tensor-core kernel: runs ONE convolution (16-bit activation-format output, so the tile kinds
of the network are exercised) once per variant of the environment switches.  Meant to be run under
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_umma --csv --log-file X python scripts/layer_probe.py SHAPE VARIANT...
SHAPE = B,H,W,cin,cout,k,stride,pad ; VARIANT = comma-separated KEY=VAL list ("-" = defaults).  Each variant launches the
layer `REPS` times (default 3); the launches appear in the CSV in variant order (the small "out" conv <*, 1, 0> follows each)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import yolo_b200  # noqa: E402

B, H, W, cin, cout, k, stride, pad = (int(v) for v in sys.argv[1].split(","))
variants = [dict(kv.split("=") for kv in a.split(",") if "=" in kv) for a in sys.argv[2:]] or [{}]
keys = sorted({k_ for v in variants for k_ in v})
spec = dict(size=[H, W], cin=cin, cout=cout, k=k, stride=stride, pad=pad, act=1, residual=int(os.environ.get("RES", "2")), bn=1)      # RES=1: + input (cin == cout)
net = yolo_b200.Net("debugconv", spec, precision=os.environ.get("PREC", "fp16x3"), max_batch=B)
rng = np.random.default_rng(0)
params = {}
for name, shape in net.param_shapes():
    leaf = name.rsplit(".", 1)[1]
    if leaf == "weight":
        params[name] = (rng.standard_normal(shape) / np.sqrt(shape[1] * shape[2] * shape[3])).astype(np.float32)
    elif leaf in ("gamma", "running_var"):
        params[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
    else:
        params[name] = rng.normal(0, 0.3, shape).astype(np.float32)
net.load_params(params)
x = torch.rand(B, 3, H, W, device="cuda")
for v in variants:
    for k_ in keys:
        os.environ.pop(k_, None)
    os.environ.update(v)
    for _ in range(int(os.environ.get("REPS", "3"))):
        net.forward(data=x)
    torch.cuda.synchronize()
print("variants:", sys.argv[2:])
