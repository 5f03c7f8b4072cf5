"""Deterministic synthetic weights + frames shared by the oracle and the CUDA path.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

The reference ships no weights (``.gitignore:10-12``); its own fallback when a weight file cannot be
loaded is ``net.collect_params().initialize(init=mxnet.init.Xavier())`` (``yolo_modules/yolo_gluon.py:194-198``).
We restate that initialiser (MXNet ``Xavier()`` defaults: rnd_type='uniform', factor_type='avg',
magnitude=3 -> U(-s, s), s = sqrt(3 / ((fan_in + fan_out) / 2))) for conv weights and follow
SURVEY.md section 8(d) for the rest: BN gamma ~ U(0.5, 1.5), beta ~ N(0, 0.1), objectness bias -4,
BN running statistics set by ONE calibration pass of the oracle in train mode on a seeded batch
(without it activations collapse through ~75 layers and every sigmoid(score) ties at 0.5).
"""
from __future__ import annotations

import numpy as np
import torch

from . import nets


def synthetic_frames(batch, size, seed=1234, layout="NCHW"):
    """uint8 HWC uniform[0,255] frames converted like ``cv_img_2_ndarray`` (yolo_gluon.py:335-357):
    transpose to CHW and divide by 255 -> fp32 (B,3,H,W).  Returns (float_nchw, uint8_nhwc)."""
    rng = np.random.default_rng(seed)
    u8 = rng.integers(0, 256, size=(batch, size[0], size[1], 3), dtype=np.uint8)
    f = (u8.astype(np.float32).transpose(0, 3, 1, 2) / np.float32(255.0)).astype(np.float32)
    return np.ascontiguousarray(f), u8


def _xavier_uniform(rng, shape):
    hw = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    fan_in, fan_out = shape[1] * hw, shape[0] * hw
    scale = np.sqrt(3.0 / ((fan_in + fan_out) / 2.0))
    return rng.uniform(-scale, scale, size=shape).astype(np.float32)


def make_params(net, spec, seed=2024, calib_batch=2, calib_size=None, calibrate=True):
    """Return ``dict[str, np.ndarray fp32]`` in canonical order (see nets.param_shapes)."""
    rng = np.random.default_rng(seed)
    shapes = nets.param_shapes(net, spec)
    p = {}
    for name, shape in shapes:
        leaf = name.rsplit(".", 1)[1]
        if leaf == "weight":
            p[name] = _xavier_uniform(rng, shape)
        elif leaf == "gamma":
            p[name] = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif leaf == "beta":
            p[name] = rng.normal(0.0, 0.1, size=shape).astype(np.float32)
        elif leaf == "running_mean":
            p[name] = np.zeros(shape, np.float32)
        elif leaf == "running_var":
            p[name] = np.ones(shape, np.float32)
        elif leaf == "bias":
            b = np.zeros(shape, np.float32)
            if name.startswith("yolo_outputs."):
                C = spec["slice_point"][-1]
                b[0::C] = -4.0                      # objectness channel of every anchor
            elif name in ("LP_branch.5.bias", "tail.conv2.bias"):
                b[0] = -4.0
            p[name] = b
        else:
            raise KeyError(name)
    if calibrate:
        size = calib_size or spec["size"]
        x, _ = synthetic_frames(calib_batch, size, seed=seed + 1)
        tp = {k: torch.from_numpy(v) for k, v in p.items()}
        with torch.no_grad():
            ctx = nets.Ctx(tp, train=True)
            # running stats := the batch statistics of the calibration batch (momentum 0 update)
            _calibrate(net, spec, ctx, torch.from_numpy(x), p)
    return p


def _calibrate(net, spec, ctx, x, p):
    import torch.nn.functional as F  # noqa: F401

    orig_bn = nets._bn

    def bn_calib(c, name, t, ch):
        mean = t.mean(dim=(0, 2, 3))
        var = t.var(dim=(0, 2, 3), unbiased=False)
        p[name + ".running_mean"] = mean.numpy().astype(np.float32).copy()
        p[name + ".running_var"] = np.maximum(var.numpy().astype(np.float32), 1e-6).copy()
        return orig_bn(c, name, t, ch)

    nets._bn = bn_calib
    try:
        nets.NET_FORWARD[net](ctx, spec, x)
    finally:
        nets._bn = orig_bn


def to_torch(params):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params.items()}


def synthetic_heads(batch, spec, seed=7, steps=None):
    """Head-only decode vectors: list of (B, H_s*W_s, A, C) fp32, shallow -> deep."""
    from . import decode
    rng = np.random.default_rng(seed)
    steps = steps or decode.init_steps(spec)
    H, W = spec["size"]
    C = spec["slice_point"][-1]
    heads = []
    for s, anchors in zip(steps, spec["all_anchors"]):
        n = (H // s) * (W // s)
        A = len(anchors)
        h = np.empty((batch, n, A, C), np.float32)
        h[..., 0] = rng.normal(-4.0, 2.0, size=(batch, n, A))
        h[..., 1:3] = rng.normal(0.0, 1.0, size=(batch, n, A, 2))
        h[..., 3:5] = rng.normal(0.0, 0.5, size=(batch, n, A, 2))
        h[..., 5:] = rng.normal(0.0, 1.0, size=(batch, n, A, C - 5))
        heads.append(h)
    return heads
