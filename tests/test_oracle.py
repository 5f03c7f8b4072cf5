"""CPU tests of the oracle: known-answer cases for the index layout / tie rules the reference pins
only through shape comments (car/YOLO.py:661-662), the committed golden fixtures, and the topology
counts of SURVEY.md section 8(d)."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import decode, nets, weights

f32 = np.float32


def _blank_heads(spec, B=1, fill=-10.0):
    steps = decode.init_steps(spec)
    H, W = spec["size"]
    C = spec["slice_point"][-1]
    return [np.full((B, (H // s) * (W // s), 3, C), fill, np.float32) for s in steps]


def test_steps_area_match_reference_comments():
    # car/YOLO.py:661-662: v1 @320x512 -> (1,640,3,30),(1,160,3,30),(1,40,3,30)
    spec = nets.spec_v1_native()
    assert decode.init_steps(spec) == [16, 32, 64]
    assert decode.init_area(spec) == [640, 160, 40]
    spec = nets.spec_dk53()
    assert decode.init_steps(spec) == [8, 16, 32]
    assert decode.init_area(spec) == [2704, 676, 169]


def test_syxhw_tables_layout():
    spec = nets.spec_micro(size=(64, 96))
    S, Y, X, Hh, Ww = decode.init_syxhw(spec)
    # scale 0: step 8, 8x12 cells; box (cell=13 -> y=1,x=1, anchor 2)
    assert S[0, 13, 2, 0] == 8 and Y[0, 13, 2, 0] == 8 and X[0, 13, 2, 0] == 8
    assert Hh[0, 13, 2, 0] == f32(0.2825) and Ww[0, 13, 2, 0] == f32(0.3456)     # anchors are (h, w)
    # first cell of scale 1 (step 16) sits right after the 96 cells of scale 0
    assert S[0, 96, 0, 0] == 16 and Y[0, 96, 0, 0] == 0 and X[0, 96 + 1, 0, 0] == 16
    assert Hh[0, 96, 1, 0] == f32(0.3703)


def test_predict_known_answer():
    """One hot box, all arithmetic checked by hand: scale 1 (step 16), cell (y=2,x=3), anchor 1."""
    spec = nets.spec_micro(size=(64, 96), C=8)
    heads = _blank_heads(spec)
    cell = 2 * 6 + 3
    row = np.array([2.0, 0.0, 0.0, 0.0, 0.0, 0.7, -1.0, 3.0], np.float32)     # ty=tx=0 -> sigmoid .5 ; th=tw=0 -> exp 1
    heads[1][0, cell, 1] = row
    pred, idx = decode.predict(spec, heads, return_index=True)
    assert idx[0] == (96 + cell) * 3 + 1
    by, bx = (0.5 * 16 + 32) / 64, (0.5 * 16 + 48) / 96
    np.testing.assert_allclose(pred[0, :5], [1 / (1 + np.exp(-2.0)), by, bx, 0.3703, 0.4351], rtol=1e-6)
    np.testing.assert_array_equal(pred[0, 5:], row[5:])


def test_predict_ties_first_index_and_saturation():
    spec = nets.spec_micro(size=(64, 96), C=8)
    heads = _blank_heads(spec)
    heads[2][0, 1, 2, 0] = 40.0      # sigmoid == 1.0f exactly
    heads[0][0, 7, 0, 0] = 20.0      # also exactly 1.0f -> lower flat index wins although the logit is smaller
    heads[1][0, 0, 0, 0] = 20.0
    _, idx = decode.predict(spec, heads, return_index=True)
    assert decode.sigmoid32(f32(20.0)) == f32(1.0)
    assert idx[0] == 7 * 3 + 0
    # exact logit ties at the scale boundaries
    heads = _blank_heads(spec)
    heads[0][0, -1, 2, 0] = 3.0      # last box of scale 0
    heads[1][0, 0, 0, 0] = 3.0       # first box of scale 1
    _, idx = decode.predict(spec, heads, return_index=True)
    assert idx[0] == 96 * 3 - 1
    heads[0][0, -1, 2, 0] = -10.0
    heads[2][0, -1, 2, 0] = 3.0      # very last box
    _, idx = decode.predict(spec, heads, return_index=True)
    assert idx[0] == 96 * 3
    heads[1][0, 0, 0, 0] = -10.0
    _, idx = decode.predict(spec, heads, return_index=True)
    assert idx[0] == (96 + 24 + 6) * 3 - 1


def test_all_equal_scores_pick_index_zero():
    spec = nets.spec_micro(size=(64, 96), C=8)
    heads = _blank_heads(spec, B=2, fill=0.0)
    _, idx = decode.predict(spec, heads, return_index=True)
    assert list(idx) == [0, 0]


def test_nms_first_kept_is_top1_and_suppression():
    spec = nets.spec_micro(size=(64, 96), C=10)
    heads = weights.synthetic_heads(4, spec, seed=3)
    _, top = decode.predict(spec, heads, return_index=True)
    res = decode.nms(spec, heads, score_thr=0.02, iou_thr=0.3, max_out=32, max_cand=512)
    for b, (rows, idx) in enumerate(res):
        assert idx[0] == top[b]
        assert np.all(np.diff(rows[:, 0]) <= 0)                 # score-descending
        assert len(set(idx.tolist())) == len(idx)
    # threshold above every score -> top-1 alone
    res = decode.nms(spec, heads, score_thr=2.0)
    for b, (rows, idx) in enumerate(res):
        assert list(idx) == [top[b]]
    # two identical boxes of the same class: the later one is suppressed; different class survives
    h = _blank_heads(spec)
    base = np.array([3.0, 0, 0, 0, 0, 0, 5, 0, 0, 0], np.float32)
    h[0][0, 10, 0] = base
    h[0][0, 10, 0, 0] = 4.0
    h[0][0, 10, 1] = base                                        # same cell, other anchor: different box -> IoU < 1
    other = base.copy(); other[6], other[7] = 0, 5
    h[1][0, 3, 0] = base
    rows, idx = decode.nms(spec, h, score_thr=0.5, iou_thr=0.99)[0]
    assert 10 * 3 in idx


def test_lp_decodes_known_answer():
    spec = nets.spec_micro(lp=True)
    lp = np.full((2, 4, 4, 10), -5.0, np.float32)
    lp[0, 2, 1] = [1.0, 0.1, -0.2, 0.3, 0.0, 10.0, -10.0, 1, 2, 3]
    lp[1, 0, 0, 0] = 0.5
    rows, idx = decode.predict_LP_batch(spec, lp, return_index=True)
    assert list(idx) == [9, 0]
    np.testing.assert_allclose(rows[0, 1:4], [100.0, -200.0, 300.0], rtol=1e-6)
    np.testing.assert_allclose(rows[0, 4:], [0.0, np.pi / 3 * (2 / (1 + np.exp(-10.0)) - 1), -np.pi / 4 * (2 / (1 + np.exp(-10.0)) - 1)], atol=1e-6)
    nchw = lp.transpose(0, 3, 1, 2).copy()
    row, i = decode.predict_LP_single(spec, nchw, return_index=True)
    assert i == 9 and row.shape == (10,)
    np.testing.assert_array_equal(row[7:], [1, 2, 3])


@pytest.mark.parametrize("name,net", [("carnet_micro", "carnet"), ("carlpnet_micro", "carlpnet"), ("lpdensenet_micro", "lpdensenet")])
def test_oracle_reproduces_golden(name, net):
    g = golden(name)
    params = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param:")}
    u8 = g["frames_u8"]
    x = torch.from_numpy((u8.astype(np.float32).transpose(0, 3, 1, 2) / f32(255)).astype(np.float32))
    if net == "lpdensenet":
        spec = nets.spec_lp_micro(); spec["size"] = [128, 128]
    else:
        spec = nets.spec_micro(size=(128, 128), lp=(net == "carlpnet"))
    with torch.no_grad():
        out = nets.forward(net, spec, params, x)
    if net == "lpdensenet":
        np.testing.assert_allclose(out.numpy(), g["out"], rtol=1e-4, atol=1e-5)
        row, idx = decode.predict_LP_single(spec, g["out"], return_index=True)
        assert idx == int(g["lp_idx"])
        np.testing.assert_allclose(row, g["lp_row"], rtol=1e-6, atol=1e-7)
        return
    heads = out if net == "carnet" else out[0]
    for i, h in enumerate(heads):
        np.testing.assert_allclose(h.numpy(), g[f"head{i}"], rtol=1e-4, atol=1e-5)
    rows, idx = decode.predict(spec, [g[f"head{i}"] for i in range(3)], return_index=True)
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_allclose(rows, g["rows"], rtol=1e-6, atol=1e-7)
    if net == "carlpnet":
        # 31 chained convs without residuals: the fp32 oracle itself is only ~1e-3 from its fp64 evaluation here
        # and moves by ~5e-4 with oneDNN's algorithm choice (DESIGN.md "numerical noise floor")
        np.testing.assert_allclose(out[1][0].numpy(), g["lp"], rtol=0, atol=5e-3)
        lrows, lidx = decode.predict_LP_batch(spec, g["lp"], return_index=True)
        np.testing.assert_array_equal(lidx, g["lp_idx"])
        np.testing.assert_allclose(lrows, g["lp_rows"], rtol=1e-6, atol=1e-7)


def test_decode_golden():
    g = golden("decode_micro")
    spec = nets.spec_micro(size=(64, 96), C=10)
    heads = [g[f"head{i}"] for i in range(3)]
    rows, idx = decode.predict(spec, heads, return_index=True)
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_array_equal(rows, g["rows"])
    assert idx[1] == 5 * 3 + 1          # saturated tie resolved to the first index (fixture construction)
    nm = decode.nms(spec, heads, score_thr=0.05, iou_thr=0.3, max_out=16, max_cand=256)
    for b, (r, i) in enumerate(nm):
        np.testing.assert_array_equal(i, g[f"nms_idx{b}"])
        np.testing.assert_array_equal(r, g[f"nms_rows{b}"])


def test_topology_counts_match_survey():
    # dk53 75 convs, v1-native 88, car_and_LP dk53 106, LPDenseNet v2 122
    def nconv(net, spec):
        return sum(1 for n, _ in nets.param_shapes(net, spec) if n.endswith(".weight"))
    assert nconv("carnet", nets.spec_dk53()) == 75
    assert nconv("carnet", nets.spec_v1_native()) == 88
    assert nconv("carlpnet", nets.spec_dk53((608, 608), 30, True)) == 106
    assert nconv("lpdensenet", nets.spec_lp_v2()) == 122


def test_head_shapes_and_order():
    spec = nets.spec_tiny()
    p = weights.to_torch(weights.make_params("carnet", spec, calibrate=False))
    x = torch.zeros(1, 3, *spec["size"])
    with torch.no_grad():
        heads = nets.forward("carnet", spec, p, x)
    assert [tuple(h.shape) for h in heads] == [(1, 96, 3, 9), (1, 24, 3, 9), (1, 6, 3, 9)]      # shallow -> deep


def test_illegal_size_is_rejected_like_reference():
    # car/v1 (6 stages) cannot run 416x416 - the concat at car/utils.py:93 fails
    spec = nets.spec_v1_native(); spec["size"] = [416, 416]
    p = weights.to_torch(weights.make_params("carnet", spec, calibrate=False))
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            nets.forward("carnet", spec, p, torch.zeros(1, 3, 416, 416))
