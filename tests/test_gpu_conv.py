"""Kernel-level parity of the convolution kernels (FFMA and tcgen05) on single layers, through the C ABI's
YOLO_NET_DEBUGCONV harness: every shape class of the networks (1x1, 3x3 s1/s2, 13x13 maps, Cout=90 heads,
ragged M tails, residual add) against a float64 evaluation of the same layer on the harness's own input."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

LEAKY, RELU = 1, 2


def run_layer(precision, B, H, W, cin, cout, k, stride, pad, act=LEAKY, residual=0, bn=1, seed=0):
    import yolo_b200
    rng = np.random.default_rng(seed)
    spec = dict(size=[H, W], cin=cin, cout=cout, k=k, stride=stride, pad=pad, act=act, residual=residual, bn=bn)
    net = yolo_b200.Net("debugconv", spec, precision=precision, max_batch=B)
    params = {}
    for name, shape in net.param_shapes():
        leaf = name.rsplit(".", 1)[1]
        if leaf == "weight":
            fan_in = shape[1] * shape[2] * shape[3]
            params[name] = (rng.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32)
        elif leaf in ("gamma", "running_var"):
            params[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        else:
            params[name] = rng.normal(0, 0.3, shape).astype(np.float32)
    net.load_params(params)
    x = rng.uniform(0, 1, size=(B, 3, H, W)).astype(np.float32)
    out = net.forward(data=torch.from_numpy(x).cuda())[0].asnumpy()           # (B, Ho, Wo, C) fp32
    pre = net.activation("pre", (B, cin, H, W))                                # exact value of the layer input
    if residual:                                                               # 1: + input; 2: activation-format output, no add
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        out = net.activation("test", (B, cout, Ho, Wo)).transpose(0, 2, 3, 1)
    t = torch.from_numpy(pre).double()
    y = F.conv2d(t, torch.from_numpy(params["test.weight"]).double(), None, stride, pad)
    if bn:
        g, b, m, v = (torch.from_numpy(params["test." + n]).double() for n in ("gamma", "beta", "running_mean", "running_var"))
        y = (y - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + 1e-5) * g[None, :, None, None] + b[None, :, None, None]
    else:
        y = y + torch.from_numpy(params["test.bias"]).double()[None, :, None, None]
    if act == LEAKY:
        y = F.leaky_relu(y, 0.1)
    elif act == RELU:
        y = F.relu(y)
    if residual == 1:
        y = y + t
    return out, y.permute(0, 2, 3, 1).numpy(), net


SHAPES = [
    # B, H, W, cin, cout, k, stride, pad, act, residual, bn
    (2, 13, 13, 64, 128, 3, 1, 1, LEAKY, 0, 1),        # odd map, ragged M tail (338 pixels)
    (3, 26, 26, 128, 64, 1, 1, 0, LEAKY, 0, 1),        # 1x1
    (2, 32, 48, 64, 128, 3, 2, 1, LEAKY, 0, 1),        # stride-2 down conv
    (1, 16, 16, 256, 90, 1, 1, 0, 0, 0, 0),            # head: Cout=90, bias, linear
    (2, 20, 20, 128, 128, 3, 1, 1, LEAKY, 1, 1),       # residual block tail
    (1, 8, 8, 512, 1024, 3, 1, 1, LEAKY, 0, 1),        # deep K = 4608, several N tiles
    (5, 7, 9, 64, 32, 1, 1, 0, LEAKY, 0, 1),           # tiny Cout, M = 315
    (1, 12, 12, 32, 64, 3, 1, 1, LEAKY, 0, 1),         # Cin % 64 != 0 -> FFMA kernel in every precision
    (2, 10, 10, 192, 48, 3, 1, 1, RELU, 0, 1),
    (2, 24, 24, 96, 64, 3, 2, 1, LEAKY, 0, 1),         # Cin % 64 == 32 -> 64-byte swizzle rows (BLOCK_K = 32)
    (1, 16, 16, 32, 32, 1, 1, 0, LEAKY, 1, 1),         # 1x1 residual, BLOCK_K = 32
    (1, 9, 9, 16, 24, 3, 1, 1, LEAKY, 0, 1),           # FFMA kernel in every precision
    (2, 26, 26, 128, 256, 3, 1, 1, LEAKY, 2, 1),       # Cout % 256 == 0, 16-bit output -> 128 x 256 tiles (merged accumulation)
    (1, 13, 13, 256, 512, 1, 1, 0, LEAKY, 2, 1),       # 1x1 on wide tiles (tiled A map), ragged M (169 pixels), two N tiles
    (2, 16, 16, 256, 256, 3, 1, 1, LEAKY, 1, 1),       # wide tiles + residual
    (2, 26, 26, 64, 256, 3, 2, 1, RELU, 2, 1),         # wide tiles, stride 2
]


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("fp16x3", 1e-5), ("bf16x6", 1e-5), ("bf16", 6e-2)])
@pytest.mark.parametrize("shape", SHAPES)
def test_single_layer(shape, precision, tol):
    B, H, W, cin, cout, k, s, p, act, res, bn = shape
    out, ref, net = run_layer(precision, B, H, W, cin, cout, k, s, p, act, res, bn, seed=hash(shape) % 1000)
    assert out.shape == ref.shape
    err = np.abs(out - ref)
    scale = max(1.0, float(np.abs(ref).max()))
    assert err.max() <= tol * scale, f"{precision} {shape}: max err {err.max():.3e} (scale {scale:.2f}), mean {err.mean():.3e}"


# The 3-channel stem (FFMA kernels: 16 x 32-pixel tiles with the input patch staged in shared memory and packed fma.f32x2 for
# Cout 16 / 32; the pixel-per-thread kernel otherwise), reached through the harness's "pre" layer, on both input contracts of
# the reference (NCHW fp32 in [0,1], cv_img_2_ndarray; uint8 HWC camera frames with the /255 fused) and on ragged tiles.
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-6), ("fp16x3", 2e-6), ("bf16x6", 2e-6)])
@pytest.mark.parametrize("u8", [False, True])
@pytest.mark.parametrize("B,H,W,cout", [(2, 32, 64, 32), (3, 23, 45, 32), (2, 17, 33, 16), (1, 40, 24, 8), (2, 16, 32, 64)])
def test_stem(B, H, W, cout, u8, precision, tol):
    import yolo_b200
    rng = np.random.default_rng(B * 1000 + H)
    spec = dict(size=[H, W], cin=cout, cout=8, k=1, stride=1, pad=0, act=0, residual=0, bn=0)
    net = yolo_b200.Net("debugconv", spec, precision=precision, max_batch=B)
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
    if u8:
        frames = rng.integers(0, 256, size=(B, H, W, 3), dtype=np.uint8)
        net.forward(data=torch.from_numpy(frames).cuda())
        x = (torch.from_numpy(frames).permute(0, 3, 1, 2).float() / 255.0).double()      # the reference divides in fp32
    else:
        xf = rng.uniform(0, 1, size=(B, 3, H, W)).astype(np.float32)
        net.forward(data=torch.from_numpy(xf).cuda())
        x = torch.from_numpy(xf).double()
    got = net.activation("pre", (B, cout, H, W))
    y = F.conv2d(x, torch.from_numpy(params["pre.weight"]).double(), None, 1, 1)
    g, b, m, v = (torch.from_numpy(params["pre." + n]).double() for n in ("gamma", "beta", "running_mean", "running_var"))
    y = (y - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + 1e-5) * g[None, :, None, None] + b[None, :, None, None]
    ref = F.leaky_relu(y, 0.1).numpy()
    scale = max(1.0, float(np.abs(ref).max()))
    err = np.abs(got - ref).max()
    assert err <= tol * scale, f"stem {precision} u8={u8} {(B, H, W, cout)}: max err {err:.3e} (scale {scale:.2f})"
    assert net.saturated() == 0
