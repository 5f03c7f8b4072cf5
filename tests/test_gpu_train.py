"""GPU parity of the training targets / losses / head gradients (yolo_loss_targets) against the oracle
(oracle/train.py, torch autograd for the gradients)."""
import numpy as np
import pytest
import torch

from oracle import decode, nets, train, weights

pytestmark = pytest.mark.gpu


def _oracle(spec, heads, labels, hp, car_rotate=False):
    targets, mask, assign = train.loss_mask(spec, labels)
    hs = [torch.from_numpy(h).requires_grad_(True) for h in heads]
    losses = train.get_loss(spec, hs, targets, mask, hp, car_rotate)
    sum(l.sum() for l in losses).backward()                       # sum(losses).backward() per device, car/YOLO.py:394
    return np.stack([l.detach().numpy() for l in losses]), assign, [h.grad.numpy() for h in hs]


@pytest.mark.parametrize("spec,B,nobj,car_rotate", [
    (nets.spec_micro(size=(64, 96), C=10), 4, 1, False),
    (nets.spec_dk53(), 8, 1, False),                       # BASELINE config 4 geometry: 10647 boxes, 24 classes
    (nets.spec_v1_native(), 6, 3, True),                   # several labels per image, rotate loss on
])
def test_loss_targets_match_oracle(spec, B, nobj, car_rotate):
    import yolo_b200
    C = spec["slice_point"][-1]
    heads = weights.synthetic_heads(B, spec, seed=5)
    labels = train.synthetic_labels(B, C - 6, nobj=nobj, seed=99, p_box=0.7)
    labels[:, :, 5] = np.where(labels[:, :, 0] >= 0, 0.3, -1.0)
    hp = dict(train.V1_HPARAMS, scale=dict(train.V1_HPARAMS["scale"], rotate=0.5))
    ol, oassign, ograd = _oracle(spec, heads, labels, hp, car_rotate)
    losses, assign, dheads = yolo_b200.loss_targets(spec, [torch.from_numpy(h).cuda() for h in heads], labels, hp["scale"],
                                                    hp["positive_weight"], hp["negative_weight"], car_rotate, with_grad=True)
    np.testing.assert_array_equal(assign.cpu().numpy(), oassign)              # bit-exact matched boxes
    np.testing.assert_allclose(losses.cpu().numpy(), ol, rtol=2e-5, atol=1e-9)
    for d, o in zip(dheads, ograd):
        np.testing.assert_allclose(d.cpu().numpy(), o, rtol=1e-4, atol=1e-9)
    assert (oassign >= 0).any()


def test_overwrite_rule_and_no_object():
    import yolo_b200
    spec = nets.spec_micro(size=(64, 96), C=10)
    heads = weights.synthetic_heads(2, spec, seed=1)
    lab = np.full((2, 3, 10), -1.0, np.float32)
    lab[0, 0] = [1, .5, .5, .3, .3, .1, 1, 0, 0, 0]
    lab[0, 2] = [2, .5, .5, .3, .3, .2, 0, 1, 0, 0]
    hp = train.V1_HPARAMS
    ol, oassign, ograd = _oracle(spec, heads, lab, hp)
    losses, assign, dheads = yolo_b200.loss_targets(spec, [torch.from_numpy(h).cuda() for h in heads], lab, hp["scale"],
                                                    hp["positive_weight"], hp["negative_weight"], False, with_grad=True)
    np.testing.assert_array_equal(assign.cpu().numpy(), oassign)
    np.testing.assert_allclose(losses.cpu().numpy(), ol, rtol=2e-5, atol=1e-9)
    for d, o in zip(dheads, ograd):
        np.testing.assert_allclose(d.cpu().numpy(), o, rtol=1e-4, atol=1e-9)


def test_driver_surface():
    import yolo_b200
    spec = dict(nets.spec_micro(size=(64, 96), C=10), classes=[0, 1, 2, 3], **train.V1_HPARAMS)
    y = yolo_b200.YOLO.__new__(yolo_b200.YOLO)
    for k, v in spec.items():
        setattr(y, k, v)
    y.spec, y.steps = spec, decode.init_steps(spec)
    heads = weights.synthetic_heads(3, spec, seed=2)
    labels = train.synthetic_labels(3, 4, seed=3, p_box=1.0)
    losses, assign = y._loss_mask_and_get_loss([torch.from_numpy(h).cuda() for h in heads], labels)
    assert losses.shape == (5, 3) and assign.shape == (3, 1)
    ol, oassign, _ = _oracle(spec, heads, labels, train.V1_HPARAMS)
    np.testing.assert_allclose(losses.cpu().numpy(), ol, rtol=2e-5, atol=1e-9)
