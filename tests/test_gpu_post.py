"""Post-decode consumers on the GPU (csrc/post.cu) against the oracle restatement / cv2, and the DenseNet-backed YOLO_dense."""
import numpy as np
import pytest
import torch

from oracle import decode, nets, post, weights

pytestmark = pytest.mark.gpu
CAM = dict(image_width=1280, image_height=720, fx=1100.0, fy=1105.0, cx=640.5, cy=355.25)


def test_azimuth_matches_cls2ang():
    import yolo_b200
    rng = np.random.default_rng(0)
    rows = rng.standard_normal((17, 30)).astype(np.float32) * 3
    rows[:, 0] = rng.uniform(0, 1, 17)
    rows[3, 6:] = 0.0                                   # uniform distribution: zero mean vector, atan2(~0, ~0)
    ang, rad = yolo_b200.cls2ang(rows)
    for b in range(17):
        a, r = post.cls2ang(rows[b, 0], rows[b, -24:])
        if b != 3:
            assert abs(((ang[b] - a + np.pi) % (2 * np.pi)) - np.pi) < 2e-6, b
        assert abs(rad[b] - r) < 2e-6


def test_plate_corners_and_unwarp_match_reference_math_and_cv2():
    import yolo_b200
    rng = np.random.default_rng(1)
    pr = yolo_b200.ProjectRectangle6D(CAM)
    poses = np.stack([[rng.uniform(-800, 800), rng.uniform(-300, 300), rng.uniform(2500, 9000), rng.uniform(-0.5, 0.5), rng.uniform(-0.7, 0.7),
                       rng.uniform(-0.4, 0.4)] for _ in range(6)]).astype(np.float32)
    got = pr(poses)
    for b in range(6):
        ref = post.project_rectangle(poses[b], CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"])
        np.testing.assert_allclose(got[b], ref, rtol=0, atol=2e-3)
    img = (rng.uniform(0, 255, size=(360, 640, 3))).astype(np.uint8)
    img[100:200, 200:400] = np.arange(200, dtype=np.uint8)[None, :, None]
    pose = np.float32([50, 20, 3000, 0.2, -0.3, 0.1])
    corners, clipped = pr.add_edges(img, pose)
    ref_c = post.project_rectangle(pose, CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"]) * np.float32([640 / 1280.0, 360 / 720.0])
    np.testing.assert_allclose(corners, ref_c, rtol=0, atol=2e-3)
    ref_img = post.add_edges(img, ref_c).astype(np.int32)
    d = np.abs(clipped.astype(np.int32) - ref_img)
    # cv2 interpolates with 1/32-pixel fixed-point coefficients: a few grey levels on a random-noise image, identical structure
    assert d.mean() < 3.0 and np.percentile(d, 99) <= 16, (d.mean(), d.max())
    assert clipped.shape == (160, 380, 3)


def test_yolo_dense_predict():
    """car/YOLO.py:864-937: CarDenseNet forward (NHWC (B, H*W, A, C)) + single-scale predict against the oracle."""
    import yolo_b200
    A, C = 5, 9
    spec = dict(nets.spec_lp_tiny(), all_anchors=[[[0.2, 0.15], [0.3, 0.4], [0.5, 0.45], [0.6, 0.7], [0.8, 0.75]]], slice_point=[1, 3, 5, 6, C],
                classes=[0, 1, 2], LP_num_class=A * C - 7)
    params = weights.make_params("lpdensenet", spec, seed=5, calib_batch=4)
    x, _ = weights.synthetic_frames(2, spec["size"], seed=7)
    y = yolo_b200.YOLO_dense(spec=spec, params=params, precision="fp32", max_batch=2)
    out = y.net.forward(data=torch.from_numpy(x).cuda())
    hs, ws = spec["size"][0] // 32, spec["size"][1] // 32
    assert out[0].shape == (2, hs * ws, A, C) and y.steps == [32]
    with torch.no_grad():
        ref = nets.forward("lpdensenet", spec, weights.to_torch(params), torch.from_numpy(x)).numpy()       # (B, A*C, hs, ws)
    ref = ref.transpose(0, 2, 3, 1).reshape(2, hs * ws, A, C)
    np.testing.assert_allclose(out[0].asnumpy(), ref, rtol=0, atol=2e-4)
    pred, idx = y.predict(out[0], return_index=True)
    opred, oidx = decode.predict(spec, [ref], steps=[32], return_index=True)
    np.testing.assert_array_equal(idx, oidx)
    np.testing.assert_allclose(pred[:, :5], opred[:, :5], rtol=0, atol=1e-4)


@pytest.mark.parametrize("src,dst", [((480, 640), (320, 512)), ((720, 1280), (416, 416)), ((1080, 1920), (608, 608)), ((320, 512), (320, 512))])
def test_resize_matches_cv2_bit_exact_when_shrinking(src, dst):
    import cv2
    import yolo_b200
    rng = np.random.default_rng(2)
    imgs = rng.integers(0, 256, size=(2,) + src + (3,), dtype=np.uint8)
    got = yolo_b200.resize_u8(imgs, dst).cpu().numpy()
    for b in range(2):
        assert np.array_equal(got[b], cv2.resize(imgs[b], (dst[1], dst[0])))


def test_resize_enlarging_is_within_one_level():
    import cv2
    import yolo_b200
    img = np.random.default_rng(3).integers(0, 256, size=(300, 400, 3), dtype=np.uint8)
    got = yolo_b200.resize_u8(img, (416, 416)).cpu().numpy()[0]
    d = np.abs(got.astype(int) - cv2.resize(img, (416, 416)).astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 2e-3
