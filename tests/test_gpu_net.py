"""GPU parity of the network forward (through the C ABI) against the oracle on shared synthetic weights.

Tolerance: the north-star bound is 1e-4 on fp32 outputs.  The reference is an fp32 computation, so the
comparison is noise-aware (tests/gpu_util.noise_aware_check): the CUDA result must be within
max(1e-4, 4 x the fp32 oracle's own distance to its fp64 evaluation) of the fp64 evaluation; selected
indices must be identical to the fp32 oracle's."""
import numpy as np
import pytest
import torch

from conftest import golden
from gpu_util import cuda_net, noise_aware_check, oracle_outputs
from oracle import decode, nets, weights

pytestmark = pytest.mark.gpu

PARITY_PRECISIONS = ["fp32", "fp16x3", "bf16x6"]     # all fp32-grade; fp16x3 / bf16x6 run on the tcgen05 tensor cores


def _run(net, spec, params, x, precision, u8=None):
    n = cuda_net(net, spec, params, precision, max_batch=x.shape[0])
    outs = n.forward(is_train=False, data=torch.from_numpy(x).cuda())
    outs[0].wait_to_read()
    res = [o.asnumpy() for o in outs]
    if u8 is not None:                                   # uint8 NHWC camera frames give the same result
        outs8 = n.forward(is_train=False, data=torch.from_numpy(u8).cuda())
        for a, b in zip(res, outs8):
            np.testing.assert_allclose(b.asnumpy(), a, rtol=0, atol=2e-5)
    return n, res


def _flatten_like(net, res):
    if net == "carnet":
        return [r.reshape(r.shape[0], -1, 3, r.shape[-1]) for r in res]
    return res


@pytest.mark.parametrize("precision", PARITY_PRECISIONS)
@pytest.mark.parametrize("name,net", [("carnet_micro", "carnet"), ("carlpnet_micro", "carlpnet"), ("lpdensenet_micro", "lpdensenet")])
def test_golden_fixture_nets(name, net, precision):
    g = golden(name)
    params = {k[6:]: g[k] for k in g.files if k.startswith("param:")}
    u8 = g["frames_u8"]
    x = (u8.astype(np.float32).transpose(0, 3, 1, 2) / np.float32(255)).astype(np.float32)
    if net == "lpdensenet":
        spec = nets.spec_lp_micro(); spec["size"] = [128, 128]
    else:
        spec = nets.spec_micro(size=(128, 128), lp=(net == "carlpnet"))
    _, res = _run(net, spec, params, x, precision, u8)
    ref64 = oracle_outputs(net, spec, params, x, torch.float64)
    if net == "lpdensenet":
        noise_aware_check(res[0], g["out"], ref64[0], what="densenet out")
        import yolo_b200
        row, idx = yolo_b200.decode_lp(torch.from_numpy(res[0]).cuda(), 1, spec["LP_r_max"])
        assert int(idx[0]) == int(g["lp_idx"])
        np.testing.assert_allclose(row[0].cpu().numpy(), g["lp_row"], rtol=0, atol=1e-4 * max(1.0, np.abs(g["lp_row"]).max()))
        return
    for i in range(3):
        noise_aware_check(res[i], g[f"head{i}"], ref64[i], what=f"head{i}")
    if net == "carlpnet":
        noise_aware_check(res[3], g["lp"], ref64[3], what="lp map")
    import yolo_b200
    rows, idx = yolo_b200.decode_top1(spec, [torch.from_numpy(r).cuda() for r in res[:3]])
    np.testing.assert_array_equal(idx.cpu().numpy(), g["idx"])
    np.testing.assert_allclose(rows.cpu().numpy()[:, :5], g["rows"][:, :5], rtol=0, atol=1e-4)


@pytest.mark.parametrize("precision", PARITY_PRECISIONS)
@pytest.mark.parametrize("net,spec,B", [
    ("carnet", nets.spec_tiny(), 3),
    ("carnet", nets.spec_tiny(size=(96, 160), C=30), 1),
    ("carlpnet", nets.spec_tiny(size=(128, 128), lp=True), 2),
    ("lpdensenet", nets.spec_lp_tiny(), 2),
])
def test_small_nets_match_oracle(net, spec, B, precision):
    params = weights.make_params(net, spec, seed=5, calib_batch=4)
    x, u8 = weights.synthetic_frames(B, spec["size"], seed=77)
    n, res = _run(net, spec, params, x, precision, u8)
    ref32 = oracle_outputs(net, spec, params, x)
    ref64 = oracle_outputs(net, spec, params, x, torch.float64)
    for i, (a, r32, r64) in enumerate(zip(res, ref32, ref64)):
        noise_aware_check(a, r32.reshape(a.shape), r64.reshape(a.shape), what=f"{net} out{i}")
    assert n.launches > 0


@pytest.mark.parametrize("precision", PARITY_PRECISIONS)
def test_layerwise_activations(precision):
    """Every named activation of the plan equals the oracle's tap (catches compensating errors)."""
    spec = nets.spec_tiny()
    params = weights.make_params("carnet", spec, seed=9, calib_batch=4)
    x, _ = weights.synthetic_frames(2, spec["size"], seed=3)
    n, _ = _run("carnet", spec, params, x, precision)
    taps = {}
    tp = weights.to_torch(params)
    with torch.no_grad():
        nets.forward("carnet", spec, tp, torch.from_numpy(x), tap=lambda k, v: taps.__setitem__(k, v.numpy().copy()))
    checked = 0
    for name, ref in taps.items():
        if name.startswith(("yolo_outputs", "transitions")) or ".body.1" in name and name.startswith("stages"):
            continue
        got = n.activation(name, ref.shape)
        np.testing.assert_allclose(got, ref, rtol=0, atol=2e-4, err_msg=name)
        checked += 1
    assert checked > 25


@pytest.mark.parametrize("precision", PARITY_PRECISIONS)
def test_dk53_416_end_to_end(precision):
    """BASELINE config 1/2 network (Darknet-53 416x416, C=30): heads, top-1 index and predict rows."""
    import yolo_b200
    spec = dict(nets.spec_dk53(), classes=list(range(24)))
    params = weights.make_params("carnet", spec, seed=2024, calib_batch=1)
    B = 2
    x, _ = weights.synthetic_frames(B, spec["size"], seed=1234)
    y = yolo_b200.YOLO(spec=spec, params=params, precision=precision, max_batch=B)
    out = y.net.forward(is_train=False, data=torch.from_numpy(x).cuda())
    out[0].wait_to_read()
    assert [o.shape for o in out] == [(B, 2704, 3, 30), (B, 676, 3, 30), (B, 169, 3, 30)]
    ref32 = oracle_outputs("carnet", spec, params, x)
    ref64 = oracle_outputs("carnet", spec, params, x, torch.float64)
    for i in range(3):
        noise_aware_check(out[i].asnumpy(), ref32[i], ref64[i], what=f"dk53 head{i}")
    pred, idx = y.predict(out, return_index=True)
    opred, oidx = decode.predict(spec, ref32, return_index=True)
    np.testing.assert_array_equal(idx, oidx)
    pred64 = decode.predict(spec, [r.astype(np.float32) for r in ref64])         # rows decoded from the fp64 heads
    assert np.abs(pred[:, :4] - opred[:, :4]).max() <= 1e-4                      # score, y, x, h vs the fp32 oracle
    head_noise = max(float(np.abs(a.astype(np.float64) - b).max()) for a, b in zip(ref32, ref64))
    noise_aware_check(pred[:, :5], opred[:, :5], pred64[:, :5], floor=max(1e-4, 4 * head_noise), what="dk53 score+bbox")   # w = exp(tw)*anchor
    sm = lambda z: np.exp(z - z.max(-1, keepdims=True)) / np.exp(z - z.max(-1, keepdims=True)).sum(-1, keepdims=True)
    np.testing.assert_allclose(sm(pred[:, 6:]), sm(opred[:, 6:]), rtol=0, atol=1e-4)   # class scores (video_node.py:246)
    # the one-call host path gives the same rows
    rows, idx2 = y.net.predict_host(torch.from_numpy(x).pin_memory())
    np.testing.assert_array_equal(idx2, oidx)
    np.testing.assert_allclose(rows, pred, rtol=0, atol=1e-6)
    assert abs(y.net.conv_flops_per_image / 1e9 - 113.263) < 1e-2


@pytest.mark.parametrize("precision", ["fp16x3"])
@pytest.mark.parametrize("name,net,spec,B", [
    ("cfg3 LPDenseNet v2 320x512", "lpdensenet", nets.spec_lp_v2(), 2),                 # licence_plate/v2/spec.yaml
    ("car/v1 native 320x512", "carnet", nets.spec_v1_native(), 1),                      # car/v1/spec.yaml (6 stages, strides 16/32/64)
    ("cfg5 car_and_LP Darknet-53 608x608", "carlpnet", nets.spec_dk53((608, 608), 30, True), 1),
])
def test_full_size_configs(name, net, spec, B, precision):
    """The other BASELINE configs at their real sizes (small batch): heads within the noise-aware bound, decoded
    selections identical to the oracle's."""
    import yolo_b200
    params = weights.make_params(net, spec, seed=77, calib_batch=1)
    x, _ = weights.synthetic_frames(B, spec["size"], seed=99)
    n, res = _run(net, spec, params, x, precision)
    ref32 = oracle_outputs(net, spec, params, x)
    ref64 = oracle_outputs(net, spec, params, x, torch.float64)
    for i, (a, r32, r64) in enumerate(zip(res, ref32, ref64)):
        noise_aware_check(a, r32.reshape(a.shape), r64.reshape(a.shape), what=f"{name} out{i}")
    if net == "lpdensenet":
        rows, idx = yolo_b200.decode_lp(torch.from_numpy(res[0]).cuda(), 1, spec["LP_r_max"])
        for b in range(B):
            orow, oi = decode.predict_LP_single(spec, ref32[0][b:b + 1], return_index=True)
            assert int(idx[b]) == oi
            np.testing.assert_allclose(rows[b].cpu().numpy(), orow, rtol=0, atol=1e-4 * max(1.0, np.abs(orow).max()))
        return
    rows, idx = yolo_b200.decode_top1(spec, [torch.from_numpy(r).cuda() for r in res[:3]])
    orows, oidx = decode.predict(spec, ref32[:3], return_index=True)
    np.testing.assert_array_equal(idx.cpu().numpy(), oidx)
    rows64 = decode.predict(spec, [r.astype(np.float32) for r in ref64[:3]])
    # decoded h, w = exp(t)*anchor carry the head logits' error 1:1, so the bound is the heads' noise-aware bound
    head_noise = max(float(np.abs(a.astype(np.float64) - b).max()) for a, b in zip(ref32[:3], ref64[:3]))
    noise_aware_check(rows.cpu().numpy()[:, :5], orows[:, :5], rows64[:, :5], floor=max(1e-4, 4 * head_noise), what=f"{name} score+bbox")
    if net == "carlpnet":
        lrows, lidx = yolo_b200.decode_lp(torch.from_numpy(res[3]).cuda(), 0, spec["LP_r_max"])
        olr, oli = decode.predict_LP_batch(spec, ref32[3], return_index=True)
        np.testing.assert_array_equal(lidx.cpu().numpy(), oli)
        assert abs(float(lrows[0, 0]) - float(olr[0, 0])) <= 1e-3      # LP branch: 31 chained convs, oracle noise ~1e-3


def test_bf16_fast_mode_reported_tolerance():
    """Single-pass bf16 (tcgen05) is the fast mode: NOT parity grade.  With synthetic (non-contractive) weights its
    error is measured and bounded loosely here; structural bugs (wrong tap / swizzle / plane) would be O(1-10)."""
    spec = nets.spec_tiny(size=(128, 128), C=30)
    params = weights.make_params("carnet", spec, seed=5, calib_batch=4)
    x, _ = weights.synthetic_frames(2, spec["size"], seed=77)
    n, res = _run("carnet", spec, params, x, "bf16")
    ref32 = oracle_outputs("carnet", spec, params, x)
    for a, r in zip(res, ref32):
        err = np.abs(a - r.reshape(a.shape))
        assert err.mean() < 0.08 and err.max() < 1.0, (err.mean(), err.max())


def test_batch_limits_and_errors():
    import yolo_b200
    spec = nets.spec_tiny()
    params = weights.make_params("carnet", spec, seed=1, calibrate=False)
    n = cuda_net("carnet", spec, params, "fp32", max_batch=2)
    with pytest.raises(yolo_b200.YoloError):
        n.forward(data=torch.zeros(3, 3, *spec["size"], device="cuda"))
    with pytest.raises(ValueError):
        n.forward(data=torch.zeros(1, 3, 32, 32, device="cuda"))
    bad = dict(params); bad.pop("stages.0.weight")
    with pytest.raises(KeyError):
        cuda_net("carnet", spec, bad)
