"""CPU tests of the training-target / loss oracle (oracle/train.py): known answers for the anchor-box layout and the
target transform (car/YOLO.py:209-240, 401-448), the overwrite rule of _loss_mask and the gluon loss definitions."""
import numpy as np
import torch

from oracle import decode, nets, train

f32 = np.float32


def test_default_ltrb_layout():
    spec = nets.spec_micro(size=(64, 96), C=10)
    ltrb = train.default_ltrb(spec)
    assert ltrb.shape == (96 + 24 + 6, 3, 4)
    # scale 0 (step 8): cell (y=1, x=2), anchor 1 (h=.2144, w=.2408): centre = ((2+.5)*8/96, (1+.5)*8/64)
    cell = 1 * 12 + 2
    l, t, r, b = ltrb[cell, 1]
    np.testing.assert_allclose([(l + r) / 2, (t + b) / 2, r - l, b - t], [2.5 * 8 / 96, 1.5 * 8 / 64, 0.2408, 0.2144], rtol=1e-6)
    # first cell of scale 2 (step 32) follows the 96 + 24 cells of the finer scales
    l, t, r, b = ltrb[120, 2]
    np.testing.assert_allclose([(l + r) / 2, (t + b) / 2], [16 / 96, 16 / 64], rtol=1e-6)


def test_find_best_known_answer():
    spec = nets.spec_micro(size=(64, 96), C=10)
    ltrb = train.default_ltrb(spec)
    # a box exactly the size of scale-1 anchor 2 (h .5708, w .4278) centred 1/4 into cell (y=1, x=3) of the 4x6 grid
    cy, cx = (1 + 0.25) * 16 / 64, (3 + 0.25) * 16 / 96
    L = np.asarray([3, cy, cx, 0.5708, 0.4278, 0.0] + [0] * 4, np.float32)
    flat, box = train.find_best(spec, ltrb, L)
    assert flat == (96 + 1 * 6 + 3) * 3 + 2
    # sigmoid(ty) = (cy - cell_centre) * H / step + 0.5 = 0.25  ->  ty = -log(1/0.25 - 1) = -log 3 ; th = tw = log 1 = 0
    np.testing.assert_allclose(box, [-np.log(3.0), -np.log(3.0), 0.0, 0.0], atol=2e-6)


def test_loss_mask_overwrite_and_absent_labels():
    spec = nets.spec_micro(size=(64, 96), C=10)
    lab = np.full((2, 3, 10), -1.0, np.float32)
    lab[0, 0] = [1, .5, .5, .3, .3, .1, 1, 0, 0, 0]
    lab[0, 2] = [2, .5, .5, .3, .3, .2, 0, 1, 0, 0]          # same box as label 0 -> overwrites it
    targets, mask, assign = train.loss_mask(spec, lab)
    assert assign[0, 0] == assign[0, 2] >= 0 and assign[0, 1] == -1 and (assign[1] == -1).all()
    assert mask.sum() == 1.0
    px, anc = divmod(int(assign[0, 0]), 3)
    assert targets[3][0, px, anc, 0] == f32(.2) and targets[4][0, px, anc, 1] == 1.0   # the LATER label's rotate / class


def test_losses_against_closed_form():
    spec = nets.spec_micro(size=(64, 96), C=10)
    heads = [np.zeros((1, n, 3, 10), np.float32) for n in (96, 24, 6)]
    lab = np.full((1, 1, 10), -1.0, np.float32)
    hp = train.V1_HPARAMS
    targets, mask, _ = train.loss_mask(spec, lab)
    losses = train.get_loss(spec, [torch.from_numpy(h) for h in heads], targets, mask, hp)
    # no object: only the score loss, log(2) * negative_weight * scale on every box
    np.testing.assert_allclose(losses[0].numpy(), [np.log(2.0) * 0.1 * 0.1], rtol=1e-6)
    for q in (1, 2, 3, 4):
        assert float(losses[q]) == 0.0
    lab[0, 0] = [0, .5, .5, .3, .3, 0, .25, .25, .25, .25]
    targets, mask, assign = train.loss_mask(spec, lab)
    losses = train.get_loss(spec, [torch.from_numpy(h) for h in heads], targets, mask, hp)
    N = 126 * 3
    t = np.concatenate([targets[1].reshape(-1, 2), targets[2].reshape(-1, 2)], 1)[int(assign[0, 0])]
    hub = lambda d: np.where(np.abs(d) > 1, np.abs(d) - .5, .5 * d * d)
    np.testing.assert_allclose(float(losses[1]), hub(t[:2]).sum() * 0.01 / (2 * N), rtol=1e-5)
    np.testing.assert_allclose(float(losses[2]), hub(t[2:]).sum() * 10.0 / (2 * N), rtol=1e-5)
    np.testing.assert_allclose(float(losses[4]), np.log(4.0) * 0.3 / N, rtol=1e-5)       # uniform logits vs uniform label
    np.testing.assert_allclose(float(losses[0]), np.log(2.0) * 0.1 * (0.1 * (N - 1) + 1.0) / N, rtol=1e-5)
