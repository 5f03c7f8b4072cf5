"""tcgen05 weight-gradient kernel (wgrad_umma.cu) against torch autograd in fp64: MN-major (transposed) operands straight from the
NHWC / [pixels][Cout] layouts, split over the pixel range with a deterministic reduction."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _wgrad(x, dz, k, stride, pad, variant=0):
    from yolo_b200 import _lib
    lib = _lib.load()
    n, h, w, cin = x.shape
    cout = dz.shape[1]
    dW = torch.full((k * k * cin, cout), float("nan"), device="cuda")
    _lib.check(lib.yolo_debug_wgrad(C.c_void_p(x.data_ptr()), C.c_void_p(dz.data_ptr()), n, h, w, cin, cout, k, stride, pad, C.c_void_p(dW.data_ptr()),
                                    variant, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return dW.cpu().numpy()


def _reference(x, dz, k, stride, pad):
    n, h, w, cin = x.shape
    cout = dz.shape[1]
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    xt = x.double().cpu().permute(0, 3, 1, 2)
    wt = torch.zeros((cout, cin, k, k), dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(xt, wt, stride=stride, padding=pad)
    y.backward(dz.double().cpu().view(n, ho, wo, cout).permute(0, 3, 1, 2))
    return wt.grad.permute(2, 3, 1, 0).reshape(k * k * cin, cout).numpy()      # [(r*k+s)*cin + c][o]


CASES = [  # n, h, w, cin, cout, k, stride, pad
    (2, 16, 16, 64, 64, 1, 1, 0),
    (2, 20, 20, 128, 256, 3, 1, 1),
    (3, 26, 26, 64, 128, 3, 2, 1),
    (4, 13, 13, 256, 512, 3, 1, 1),
    (1, 13, 13, 512, 256, 1, 1, 0),
    (16, 52, 52, 128, 256, 3, 1, 1),      # a Darknet-53 layer at the training batch: splits + many pixel blocks
    (2, 24, 24, 32, 64, 3, 2, 1),         # Cin = 32, plane-interleaved activations: the transposed kernel (stages.1.0)
    (3, 20, 28, 32, 64, 3, 1, 1),         # ... stride 1 (stages.1.1.body.1)
    (2, 16, 16, 32, 128, 1, 1, 0),        # ... two real dz blocks per tile, single tap
]


@pytest.mark.parametrize("case", CASES)
def test_wgrad_matches_autograd(case):
    n, h, w, cin, cout, k, stride, pad = case
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn((n, h, w, cin), device="cuda", generator=g)
    x = torch.where(x > 0, x, 0.1 * x)                                         # leaky-ReLU-like activations
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    dz = torch.randn((n * ho * wo, cout), device="cuda", generator=g)
    ref = _reference(x, dz, k, stride, pad)
    got = _wgrad(x, dz, k, stride, pad, 0)
    scale = np.abs(ref).max()
    err = np.abs(got - ref).max() / scale
    if not err < 2e-6:
        alt = _wgrad(x, dz, k, stride, pad, 1)
        err1 = np.abs(alt - ref).max() / scale
        raise AssertionError(f"wgrad {case}: rel err {err:.3e} (swapped LBO/SBO descriptor convention: {err1:.3e})")
    rel_l2 = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert rel_l2 < 1e-6, rel_l2
    assert np.array_equal(got, _wgrad(x, dz, k, stride, pad, 0))              # deterministic: no atomics
