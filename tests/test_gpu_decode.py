"""GPU parity of the fused decode + selection kernel against the oracle (bit-exact indices; rows are
compared bit-for-bit as well because the kernel evaluates the oracle's fp32 expression order)."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import decode, nets, weights

pytestmark = pytest.mark.gpu


def _cuda(heads):
    return [torch.from_numpy(h).cuda() for h in heads]


def _top1(spec, heads):
    import yolo_b200
    rows, idx = yolo_b200.decode_top1(spec, _cuda(heads))
    return rows.cpu().numpy(), idx.cpu().numpy()


@pytest.mark.parametrize("spec,B,seed", [
    (nets.spec_micro(size=(64, 96), C=10), 3, 7),
    (nets.spec_v1_native(), 5, 8),          # 2520 boxes (car/YOLO.py:661-662 shapes)
    (nets.spec_dk53(), 4, 9),               # 10647 boxes
    (nets.spec_dk53((608, 608), 80), 2, 10),
    (nets.spec_tiny(C=7), 1, 11),           # a single class channel
])
def test_top1_matches_oracle(spec, B, seed):
    heads = weights.synthetic_heads(B, spec, seed=seed)
    rows, idx = _top1(spec, heads)
    orows, oidx = decode.predict(spec, heads, return_index=True)
    np.testing.assert_array_equal(idx, oidx)
    np.testing.assert_allclose(rows, orows, rtol=0, atol=1e-6)
    assert np.array_equal(rows, orows), "rows are expected to be bit-identical to the oracle"


def test_top1_golden_fixture():
    g = golden("decode_micro")
    spec = nets.spec_micro(size=(64, 96), C=10)
    heads = [g[f"head{i}"] for i in range(3)]
    rows, idx = _top1(spec, heads)
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_array_equal(rows, g["rows"])


def test_top1_adversarial_ties_saturation_boundaries():
    spec = nets.spec_dk53()
    rng = np.random.default_rng(5)
    heads = weights.synthetic_heads(6, spec, seed=21)
    n0, n1, n2 = (h.shape[1] for h in heads)
    heads[0][0, :, :, 0] = 0.0; heads[1][0, :, :, 0] = 0.0; heads[2][0, :, :, 0] = 0.0       # all tie -> index 0
    heads[2][1, n2 - 1, 2, 0] = 50.0                                                         # very last box
    heads[1][2, 0, 0, 0] = 18.0; heads[0][2, n0 - 1, 2, 0] = 17.5                            # both saturate to 1.0f: lower index wins
    heads[0][3, 100, 1, 0] = 12.345; heads[1][3, 7, 2, 0] = 12.345                           # exact logit tie across scales
    for h in heads:                                                                          # extreme box logits
        h[4, :, :, 3:5] = rng.choice([-10.0, 10.0], size=h[4, :, :, 3:5].shape)
    heads[0][5, :, :, 0] = rng.uniform(16.0, 18.0, size=heads[0][5, :, :, 0].shape)          # the saturation boundary zone
    rows, idx = _top1(spec, heads)
    orows, oidx = decode.predict(spec, heads, return_index=True)
    np.testing.assert_array_equal(idx, oidx)
    assert oidx[0] == 0 and oidx[1] == (n0 + n1 + n2) * 3 - 1 and oidx[2] == n0 * 3 - 1 and oidx[3] == 100 * 3 + 1
    assert np.array_equal(rows, orows)


def test_top1_full_size_property():
    """BASELINE config 2 size (B=32, 10647 boxes): planted maxima are found, and the result is invariant to
    permuting the batch (size-independent properties; the oracle also runs here in a few seconds)."""
    spec = nets.spec_dk53()
    B = 32
    heads = weights.synthetic_heads(B, spec, seed=33)
    rng = np.random.default_rng(1)
    total = sum(h.shape[1] * 3 for h in heads)
    planted = rng.integers(0, total, size=B)
    offs = np.cumsum([0] + [h.shape[1] * 3 for h in heads])
    for b, j in enumerate(planted):
        s = int(np.searchsorted(offs, j, side="right") - 1)
        loc = j - offs[s]
        heads[s][b, loc // 3, loc % 3, 0] = 15.0
    rows, idx = _top1(spec, heads)
    np.testing.assert_array_equal(idx, planted.astype(np.int32))
    perm = rng.permutation(B)
    rows_p, idx_p = _top1(spec, [h[perm] for h in heads])
    np.testing.assert_array_equal(idx_p, idx[perm])
    assert np.array_equal(rows_p, rows[perm])
    orows, oidx = decode.predict(spec, heads, return_index=True)
    np.testing.assert_array_equal(idx, oidx)
    assert np.array_equal(rows, orows)


@pytest.mark.parametrize("spec,B,thr,iou,max_out,max_cand", [
    (nets.spec_micro(size=(64, 96), C=10), 3, 0.05, 0.3, 16, 256),
    (nets.spec_dk53(), 3, 0.3, 0.45, 100, 1024),
    (nets.spec_dk53(), 2, 0.1, 0.45, 50, 1024),            # ~1900 raw candidates -> truncated to max_cand after the sort
    (nets.spec_dk53(), 2, 0.2, 0.1, 8, 64),
    (nets.spec_v1_native(), 2, 0.9999, 0.5, 10, 100),      # nothing passes -> top-1 alone
])
def test_nms_matches_oracle(spec, B, thr, iou, max_out, max_cand):
    import yolo_b200
    heads = weights.synthetic_heads(B, spec, seed=17)
    rows, idx, cnt = yolo_b200.decode_nms(spec, _cuda(heads), thr, iou, max_out, max_cand)
    rows, idx, cnt = rows.cpu().numpy(), idx.cpu().numpy(), cnt.cpu().numpy()
    ref = decode.nms(spec, heads, thr, iou, max_out, max_cand)
    _, top = decode.predict(spec, heads, return_index=True)
    for b, (orows, oidx) in enumerate(ref):
        assert cnt[b] == len(oidx)
        np.testing.assert_array_equal(idx[b, :cnt[b]], oidx)
        assert np.array_equal(rows[b, :cnt[b]], orows)
        assert idx[b, 0] == top[b]                      # consistency with the reference's top-1


def test_nms_golden_and_overflow_flag():
    import yolo_b200
    g = golden("decode_micro")
    spec = nets.spec_micro(size=(64, 96), C=10)
    heads = [g[f"head{i}"] for i in range(3)]
    rows, idx, cnt = yolo_b200.decode_nms(spec, _cuda(heads), 0.05, 0.3, 16, 256)
    for b in range(3):
        n = int(cnt[b])
        np.testing.assert_array_equal(idx[b, :n].cpu().numpy(), g[f"nms_idx{b}"])
        assert np.array_equal(rows[b, :n].cpu().numpy(), g[f"nms_rows{b}"])
    spec = nets.spec_dk53()
    heads = weights.synthetic_heads(1, spec, seed=2)
    for h in heads:
        h[..., 0] = 3.0                                 # 10647 candidates > 4096 -> overflow is reported, not mis-ordered
    _, _, cnt = yolo_b200.decode_nms(spec, _cuda(heads), 0.5, 0.5, 10, 100)
    assert int(cnt[0]) == -10647


def test_lp_decodes_match_oracle():
    import yolo_b200
    spec = nets.spec_dk53((608, 608), 30, True)
    rng = np.random.default_rng(4)
    lp = rng.normal(0, 2, size=(5, 76, 76, 10)).astype(np.float32)
    lp[1, 0, 0, 0] = 30.0; lp[1, 5, 5, 0] = 40.0          # saturation tie -> first
    rows, idx = yolo_b200.decode_lp(torch.from_numpy(lp).cuda(), 0, spec["LP_r_max"])
    orows, oidx = decode.predict_LP_batch(spec, lp, return_index=True)
    np.testing.assert_array_equal(idx.cpu().numpy(), oidx)
    np.testing.assert_allclose(rows.cpu().numpy(), orows, rtol=1e-6, atol=1e-7)
    spec = nets.spec_lp_v2()
    out = rng.normal(0, 2, size=(3, 10, 10, 16)).astype(np.float32)
    rows, idx = yolo_b200.decode_lp(torch.from_numpy(out).cuda(), 1, spec["LP_r_max"])
    for b in range(3):
        orow, oi = decode.predict_LP_single(spec, out[b:b + 1], return_index=True)
        assert int(idx[b]) == oi
        np.testing.assert_allclose(rows[b].cpu().numpy(), orow, rtol=1e-6, atol=1e-7)


def test_driver_predict_surface():
    """The reference-facing call: YOLO.predict(list of heads) -> np.float32 (B, 6+num_class)."""
    import yolo_b200
    spec = dict(nets.spec_dk53(), classes=list(range(24)))
    y = yolo_b200.YOLO.__new__(yolo_b200.YOLO)              # decode needs no network
    y.spec, y.steps = spec, decode.init_steps(spec)
    heads = weights.synthetic_heads(2, spec, seed=3)
    out = y.predict([yolo_b200.NDArray(t) for t in _cuda(heads)])
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == (2, 30)
    assert np.array_equal(out, decode.predict(spec, heads))
    rows, idx = y.predict(_cuda(heads), return_index=True)
    np.testing.assert_array_equal(idx, decode.predict(spec, heads, return_index=True)[1])


def test_empty_batch_and_ragged_batch():
    """Empty input (batch 0) is a no-op with status OK; a batch smaller than max_batch runs on the same handle."""
    import ctypes as C
    import yolo_b200
    from yolo_b200 import _lib, api
    spec = nets.spec_micro(size=(64, 96), C=10)
    lib = _lib.load()
    g = api.make_geom(spec)
    heads = [torch.zeros((0, n, 3, 10), device="cuda") for n in (96, 24, 6)]
    ptrs = (C.c_void_p * 3)(*[1, 1, 1])          # never dereferenced for batch 0
    rows = torch.zeros((1, 10), device="cuda"); idx = torch.zeros((1,), dtype=torch.int32, device="cuda")
    assert lib.yolo_decode_top1(C.byref(g), ptrs, 0, C.c_void_p(rows.data_ptr()), C.c_void_p(idx.data_ptr()), None) == 0
    assert lib.yolo_decode_lp(C.c_void_p(rows.data_ptr()), 0, 4, 4, 10, 0, C.byref((C.c_float * 3)(45, 60, 45)), C.c_void_p(rows.data_ptr()), None, None) == 0
    r, i = yolo_b200.decode_top1(spec, [torch.from_numpy(h).cuda() for h in weights.synthetic_heads(1, spec, seed=4)])
    assert r.shape == (1, 10) and i.shape == (1,)
    # ragged: handle built for 4 images, called with 1 and 3
    params = weights.make_params("carnet", spec, seed=1, calib_batch=2)
    net = yolo_b200.Net("carnet", spec, precision="fp16x3", max_batch=4).load_params(params)
    x, _ = weights.synthetic_frames(4, spec["size"], seed=2)
    full = [o.asnumpy() for o in net.forward(data=torch.from_numpy(x).cuda())]
    for b in (1, 3):
        part = [o.asnumpy() for o in net.forward(data=torch.from_numpy(x[:b]).cuda())]
        for a, f in zip(part, full):
            np.testing.assert_array_equal(a, f[:b])          # per-image results do not depend on the batch


def test_maximum_boxes_608_c80():
    """Largest head geometry of the BASELINE configs (608x608, C=80: 22743 boxes, 7.3 MB per image): planted maximum found."""
    spec = nets.spec_dk53((608, 608), 80)
    heads = weights.synthetic_heads(2, spec, seed=12)
    heads[2][1, -1, 2, 0] = 30.0
    rows, idx = _top1(spec, heads)
    orows, oidx = decode.predict(spec, heads, return_index=True)
    np.testing.assert_array_equal(idx, oidx)
    assert idx[1] == 22743 - 1 and np.array_equal(rows, orows)
