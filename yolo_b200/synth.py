"""Random-init parameters for benchmarking / smoke runs (no checkpoints exist offline).

Mirrors the reference's own fallback when no weight file loads: ``initialize(init=mxnet.init.Xavier())``
(yolo_modules/yolo_gluon.py:194-198) - uniform, factor_type 'avg', magnitude 3 - with BN gamma chosen so
that activations keep O(1) scale through the stack (running_mean 0, running_var 1).
"""
from __future__ import annotations

import numpy as np


def random_params(param_shapes, seed=0, obj_bias=-4.0, channels_per_anchor=None):
    rng = np.random.default_rng(seed)
    p = {}
    for name, shape in param_shapes:
        leaf = name.rsplit(".", 1)[1]
        if leaf == "weight":
            hw = int(np.prod(shape[2:]))
            fan_in, fan_out = shape[1] * hw, shape[0] * hw
            s = np.sqrt(3.0 / ((fan_in + fan_out) / 2.0))
            # variance-preserving gain folded into the weights: Xavier('avg') shrinks by 2*fan_in/(fan_in+fan_out)
            g = np.sqrt((fan_in + fan_out) / (2.0 * fan_in)) * 1.3
            p[name] = (rng.uniform(-s, s, size=shape) * g).astype(np.float32)
        elif leaf == "gamma":
            p[name] = rng.uniform(0.8, 1.2, size=shape).astype(np.float32)
        elif leaf == "beta":
            p[name] = rng.normal(0.0, 0.1, size=shape).astype(np.float32)
        elif leaf == "running_mean":
            p[name] = np.zeros(shape, np.float32)
        elif leaf == "running_var":
            p[name] = np.ones(shape, np.float32)
        elif leaf == "bias":
            b = np.zeros(shape, np.float32)
            if channels_per_anchor and name.startswith("yolo_outputs."):
                b[0::channels_per_anchor] = obj_bias
            p[name] = b
        else:
            raise KeyError(name)
    return p


def calibrated_params(net, frames, seed=0, channels_per_anchor=None):
    """Random-init parameters whose BatchNorm running statistics are the batch statistics of ``frames`` on this very network -
    one train-mode forward on the GPU (Trainer.calibrate_bn).  Without it 23 residual blocks with running_var = 1 let the
    activations grow by orders of magnitude and every sigmoid saturates.  Loads them into ``net`` and returns the dict."""
    from .api import Trainer
    p = random_params(net.param_shapes(), seed=seed, channels_per_anchor=channels_per_anchor)
    net.load_params(p)
    tr = Trainer(net, join_nccl=False)
    tr.calibrate_bn(frames)
    for name, shape in net.param_shapes():
        if name.endswith((".running_mean", ".running_var")):
            p[name] = tr.get_param(name, shape)
    net.load_params(p)                   # also releases the trainer (its arena is several GB at large batch)
    net._trainer = None
    return p
