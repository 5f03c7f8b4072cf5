"""Oracle restatement (torch-CPU fp32, functional) of the reference network topologies.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

Follows (file:line under /root/reference):
  * ``yolo_modules/basic_yolo.py:8-39``   BasicYOLONet.__init__  (stem, stages, pyramid)
  * ``yolo_modules/basic_yolo.py:91-105`` YOLOOutput (1x1 conv + bias, transpose(0,2,3,1), reshape)
  * ``yolo_modules/basic_yolo.py:108-123`` YOLOPyrmaid (block channel = the pyramid channel itself)
  * ``car/utils.py:68-95``                CarNet.hybrid_forward (heads returned shallow -> deep)
  * ``car_and_LP/YOLO.py:47-95``          CarLPNet (+ 5 chained detection blocks and a 1x1 conv)
  * ``licence_plate/LP_detection.py:59-97`` LPDenseNet
and the gluoncv==0.4.0b20181129 block definitions the reference imports (not vendored; restated
from the pinned release):
  ``_conv2d`` = Conv2D(no bias) -> BatchNorm(eps 1e-5, momentum 0.9) -> LeakyReLU(0.1);
  ``DarknetBasicBlockV3(c)`` = x + [_conv2d(c,1,0,1), _conv2d(2c,3,1,1)](x);
  ``YOLODetectionBlockV3(c)``: body = [_conv2d(c,1), _conv2d(2c,3)] x2 + _conv2d(c,1); tip = _conv2d(2c,3);
  ``_upsample`` = repeat x2 on W then H (nearest);
  DenseNet ``_make_dense_layer`` / ``_make_transition``.

Parameters are a flat ``dict[str, torch.Tensor]`` keyed by hierarchical names mirroring the gluon
attribute names (stages / transitions / yolo_blocks / yolo_outputs / LP_branch); the canonical
enumeration order is first-use order of the forward pass.
Every function takes a ``Ctx`` which either *collects* (name, shape) pairs (dry run) or *reads*
tensors, so the enumeration and the forward can never disagree.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # gluon BatchNorm() default, used by gluoncv _conv2d
BN_MOMENTUM = 0.9      # gluon BatchNorm() default
LEAKY = 0.1            # gluoncv _conv2d: LeakyReLU(0.1)


class Ctx:
    """Parameter access context: collect shapes (params=None) or read tensors."""

    def __init__(self, params=None, train=False):
        self.params = params
        self.train = train
        self.shapes = []          # [(name, shape)] in first-use order
        self.new_stats = {}       # train mode: updated running stats (name -> tensor)
        self.tap = None           # optional callable(name, tensor) to record activations

    def get(self, name, shape):
        if self.params is None:
            self.shapes.append((name, tuple(shape)))
            return torch.zeros(shape)
        t = self.params[name]
        assert tuple(t.shape) == tuple(shape), (name, tuple(t.shape), tuple(shape))
        return t


def _bn(ctx, name, x, c):
    g = ctx.get(name + ".gamma", (c,))
    b = ctx.get(name + ".beta", (c,))
    m = ctx.get(name + ".running_mean", (c,))
    v = ctx.get(name + ".running_var", (c,))
    if ctx.train:
        # batch statistics per device (num_sync_bn_devices=-1, car/YOLO.py:94-96)
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        ctx.new_stats[name + ".running_mean"] = (m * BN_MOMENTUM + mean.detach() * (1 - BN_MOMENTUM))
        ctx.new_stats[name + ".running_var"] = (v * BN_MOMENTUM + var.detach() * (1 - BN_MOMENTUM))
        return F.batch_norm(x, None, None, g, b, True, 0.0, BN_EPS)
    return F.batch_norm(x, m, v, g, b, False, 0.0, BN_EPS)


def conv2d_bn_leaky(ctx, name, x, cout, k, pad, stride):
    """gluoncv ``_conv2d(channel, kernel, padding, stride)``."""
    cin = x.shape[1]
    w = ctx.get(name + ".weight", (cout, cin, k, k))
    y = F.conv2d(x, w, None, stride, pad)
    y = _bn(ctx, name, y, cout)
    y = F.leaky_relu(y, LEAKY)
    if ctx.tap is not None:
        ctx.tap(name, y)
    return y


def conv2d_bias(ctx, name, x, cout, k, pad=0):
    """gluon.nn.Conv2D(cout, kernel_size=k, padding=pad) with bias, no activation."""
    cin = x.shape[1]
    w = ctx.get(name + ".weight", (cout, cin, k, k))
    b = ctx.get(name + ".bias", (cout,))
    y = F.conv2d(x, w, b, 1, pad)
    if ctx.tap is not None:
        ctx.tap(name, y)
    return y


def darknet_basic_block(ctx, name, x, c):
    y = conv2d_bn_leaky(ctx, name + ".body.0", x, c, 1, 0, 1)
    y = conv2d_bn_leaky(ctx, name + ".body.1", y, c * 2, 3, 1, 1)
    y = x + y
    if ctx.tap is not None:
        ctx.tap(name, y)
    return y


def detection_block(ctx, name, x, c):
    """gluoncv YOLODetectionBlockV3(c) -> (route, tip)."""
    r = x
    for i in range(2):
        r = conv2d_bn_leaky(ctx, f"{name}.body.{2 * i}", r, c, 1, 0, 1)
        r = conv2d_bn_leaky(ctx, f"{name}.body.{2 * i + 1}", r, c * 2, 3, 1, 1)
    r = conv2d_bn_leaky(ctx, f"{name}.body.4", r, c, 1, 0, 1)
    tip = conv2d_bn_leaky(ctx, f"{name}.tip", r, c * 2, 3, 1, 1)
    return r, tip


def upsample2(x):
    """gluoncv ``_upsample(x, stride=2)``: repeat on the last axis, then on the one before."""
    return x.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)


def yolo_output(ctx, name, x, num_anchors, channel):
    """YOLOOutput (basic_yolo.py:91-105): (B,A*C,H,W) -> (B,H*W,A,C)."""
    y = conv2d_bias(ctx, name, x, channel * num_anchors, 1)
    y = y.permute(0, 2, 3, 1)
    return y.reshape(y.shape[0], -1, num_anchors, channel)


def _backbone(ctx, spec, x):
    """BasicYOLONet stages (basic_yolo.py:18-27) -> routes of the last n_pyramid stages."""
    layers, channels = spec["layers"], spec["channels"]
    assert len(layers) == len(channels) - 1
    npyr = len(spec["all_anchors"])
    x = conv2d_bn_leaky(ctx, "stages.0", x, channels[0], 3, 1, 1)
    routes = []
    nstage = len(layers) + 1
    for s, (nlayer, channel) in enumerate(zip(layers, channels[1:]), start=1):
        x = conv2d_bn_leaky(ctx, f"stages.{s}.0", x, channel, 3, 1, 2)
        for j in range(1, nlayer + 1):
            x = darknet_basic_block(ctx, f"stages.{s}.{j}", x, channel // 2)
        if s >= nstage - npyr:
            routes.append(x)
    return x, routes


def carnet_forward(ctx, spec, x, lp_branch=False):
    """CarNet.hybrid_forward (car/utils.py:68-95); with ``lp_branch`` CarLPNet (car_and_LP/YOLO.py:62-95).

    Returns heads ordered shallow -> deep, each (B, H_s*W_s, A, C); with lp_branch also
    LP_output (B, H_s, W_s, LP_slice_point[-1]) computed on the shallowest concat map.
    """
    anchors = spec["all_anchors"]
    npyr = len(anchors)
    pyr_channels = spec["channels"][-npyr:][::-1]
    C = spec["slice_point"][-1]
    x, routes = _backbone(ctx, spec, x)
    outs = []
    lp_out = None
    for i in range(npyr):
        last = i >= npyr - 1
        if last and lp_branch:
            lpc = spec["channels"][-3]
            t = x
            for b in range(5):
                _, t = detection_block(ctx, f"LP_branch.{b}", t, lpc)
            t = conv2d_bias(ctx, "LP_branch.5", t, spec["LP_slice_point"][-1], 1)
            lp_out = t.permute(0, 2, 3, 1)
        x, tip = detection_block(ctx, f"yolo_blocks.{i}", x, pyr_channels[i])
        outs.append(yolo_output(ctx, f"yolo_outputs.{i}", tip, len(anchors[::-1][i]), C))
        if last:
            break
        x = conv2d_bn_leaky(ctx, f"transitions.{i}", x, pyr_channels[i + 1], 1, 0, 1)
        x = torch.cat([upsample2(x), routes[::-1][i + 1]], dim=1)
    heads = outs[::-1]
    if lp_branch:
        return heads, [lp_out]
    return heads


def _bn_relu(ctx, name, x):
    return F.relu(_bn(ctx, name, x, x.shape[1]))


def lpdensenet_forward(ctx, spec, x):
    """LPDenseNet (licence_plate/LP_detection.py:59-97) -> NCHW (B, 7+LP_num_class, H/32, W/32)."""
    nif, g, cfg = spec["num_init_features"], spec["growth_rate"], spec["block_config"]
    bn_size = spec.get("bn_size", 4)
    w = ctx.get("stem.conv.weight", (nif, x.shape[1], 7, 7))
    x = F.conv2d(x, w, None, 2, 3)
    x = _bn_relu(ctx, "stem.bn", x)
    x = F.max_pool2d(x, 3, 2, 1)
    nf = nif
    for b, nl in enumerate(cfg, start=1):
        for l in range(nl):
            p = f"block{b}.layer{l}"
            y = _bn_relu(ctx, p + ".bn1", x)
            y = F.conv2d(y, ctx.get(p + ".conv1.weight", (bn_size * g, y.shape[1], 1, 1)))
            y = _bn_relu(ctx, p + ".bn2", y)
            y = F.conv2d(y, ctx.get(p + ".conv2.weight", (g, y.shape[1], 3, 3)), None, 1, 1)
            x = torch.cat([x, y], dim=1)
            if ctx.tap is not None:
                ctx.tap(p, x)
        nf = nf + nl * g
        if b != len(cfg):
            p = f"trans{b}"
            y = _bn_relu(ctx, p + ".bn", x)
            y = F.conv2d(y, ctx.get(p + ".conv.weight", (nf // 2, nf, 1, 1)))
            x = F.avg_pool2d(y, 2, 2)
            nf = nf // 2
            if ctx.tap is not None:
                ctx.tap(p, x)
    x = _bn_relu(ctx, "tail.bn1", x)
    x = F.conv2d(x, ctx.get("tail.conv1.weight", (512, nf, 3, 3)), ctx.get("tail.conv1.bias", (512,)), 1, 1)
    x = _bn_relu(ctx, "tail.bn2", x)
    nout = 7 + spec["LP_num_class"]
    x = F.conv2d(x, ctx.get("tail.conv2.weight", (nout, 512, 1, 1)), ctx.get("tail.conv2.bias", (nout,)))
    return x


NET_FORWARD = {
    "carnet": lambda ctx, spec, x: carnet_forward(ctx, spec, x, False),
    "carlpnet": lambda ctx, spec, x: carnet_forward(ctx, spec, x, True),
    "lpdensenet": lpdensenet_forward,
}


def param_shapes(net, spec):
    """Canonical [(name, shape)] list, obtained by a dry run on a minimal-size zero image."""
    ctx = Ctx(None)
    h = w = 64
    if net != "lpdensenet":
        h = w = 2 ** (len(spec["layers"]) + 1)
    with torch.no_grad():
        NET_FORWARD[net](ctx, spec, torch.zeros(1, 3, h, w))
    seen, out = set(), []
    for n, s in ctx.shapes:
        if n not in seen:
            seen.add(n)
            out.append((n, s))
    return out


def forward(net, spec, params, x, train=False, tap=None):
    """Run the oracle.  ``x``: (B,3,H,W) fp32 in [0,1] (yolo_gluon.py:335-357).  Returns torch tensors."""
    ctx = Ctx(params, train)
    ctx.tap = tap
    out = NET_FORWARD[net](ctx, spec, x)
    if train:
        return out, ctx.new_stats
    return out


# ---------------------------------------------------------------------------------------------
# Specs (the reference's own spec.yaml schema, car/v1/spec.yaml:1-41)
# ---------------------------------------------------------------------------------------------
V1_ANCHORS = [
    [[0.2216, 0.1552], [0.2144, 0.2408], [0.2825, 0.3456]],
    [[0.3959, 0.2706], [0.3703, 0.4351], [0.5708, 0.4278]],
    [[0.4345, 0.6063], [0.5584, 0.7174], [0.7448, 0.6772]]]


def spec_dk53(size=(416, 416), C=30, lp=False):
    """Canonical Darknet-53 expressed in the reference's schema."""
    s = dict(size=list(size), layers=[1, 2, 8, 8, 4], channels=[32, 64, 128, 256, 512, 1024],
             slice_point=[1, 3, 5, 6, C], all_anchors=V1_ANCHORS, use_fp16=False)
    if lp:
        s.update(LP_slice_point=[1, 3, 4, 7, 10], LP_r_max=[45, 60, 45], LP_num_class=3)
    return s


def spec_v1_native(C=30, lp=False):
    """car/v1/spec.yaml:3-14 (6 stages, strides 16/32/64, 320x512)."""
    s = dict(size=[320, 512], layers=[1, 4, 4, 8, 8, 4], channels=[16, 32, 64, 128, 256, 512, 1024],
             slice_point=[1, 3, 5, 6, C], all_anchors=V1_ANCHORS, use_fp16=False)
    if lp:
        s.update(LP_slice_point=[1, 3, 4, 7, 10], LP_r_max=[45, 60, 45], LP_num_class=3)
    return s


def spec_tiny(size=(64, 96), C=9, lp=False):
    """Small spec for fast tests (same schema; strides 8/16/32)."""
    s = dict(size=list(size), layers=[1, 1, 2, 1, 1], channels=[8, 16, 32, 64, 128, 256],
             slice_point=[1, 3, 5, 6, C], all_anchors=V1_ANCHORS, use_fp16=False)
    if lp:
        s.update(LP_slice_point=[1, 3, 4, 7, 10], LP_r_max=[45, 60, 45], LP_num_class=3)
    return s


def spec_micro(size=(64, 64), C=8, lp=False):
    """Very small spec whose parameters fit in a committed fixture (tests/golden)."""
    s = dict(size=list(size), layers=[1, 1, 1, 1, 1], channels=[4, 8, 8, 16, 16, 32],
             slice_point=[1, 3, 5, 6, C], all_anchors=V1_ANCHORS, use_fp16=False)
    if lp:
        s.update(LP_slice_point=[1, 3, 4, 7, 10], LP_r_max=[45, 60, 45], LP_num_class=3)
    return s


def spec_lp_v2():
    """licence_plate/v2/spec.yaml:1-10."""
    return dict(size=[320, 512], LP_slice_point=[1, 3, 4, 7, 10], LP_r_max=[45, 60, 45], LP_num_class=3,
                num_init_features=64, growth_rate=16, block_config=[6, 12, 24, 16])


def spec_lp_micro():
    return dict(size=[64, 64], LP_slice_point=[1, 3, 4, 7, 10], LP_r_max=[45, 60, 45], LP_num_class=3,
                num_init_features=8, growth_rate=4, block_config=[2, 2, 2, 2])


def spec_lp_tiny():
    return dict(size=[64, 96], LP_slice_point=[1, 3, 4, 7, 10], LP_r_max=[45, 60, 45], LP_num_class=3,
                num_init_features=16, growth_rate=8, block_config=[2, 3, 2, 2])
