"""Oracle restatement (numpy fp32 / torch) of the reference's training targets and losses.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

Follows (file:line under /root/reference):
  * ``car/YOLO.py:209-240``   _get_default_ltrb   (anchor boxes at the cell centres, normalised ltrb)
  * ``yolo_modules/yolo_gluon.py:127-168`` get_iou (mode 2: target = [c, y, x, h, w])
  * ``car/YOLO.py:401-448``   _find_best          (argmax IoU -> (pixel, anchor), ty/tx/th/tw targets)
  * ``car/YOLO.py:450-480``   _loss_mask          (dense targets; later labels overwrite earlier ones)
  * ``car/YOLO.py:482-489``   _score_weight
  * ``car/YOLO.py:491-498``   _get_loss           (rotate weight 0 unless car_rotate)
and the gluon losses it instantiates at ``car/YOLO.py:185-190`` (mxnet gluon/loss.py, restated):
  LogisticLoss(label_format='binary'): l = 2*label-1; relu(-p*l) + softrelu(-|p*l|);
  HuberLoss(rho=1): |d| > 1 ? |d| - 0.5 : 0.5*d^2;
  SoftmaxCrossEntropyLoss(from_logits=False, sparse_label=False): -sum(log_softmax(p) * label, -1, keepdims);
  each multiplied by its sample weight and averaged over all non-batch axes -> (B,).
"""
from __future__ import annotations

import numpy as np
import torch

from . import decode

f32 = np.float32


def default_ltrb(spec, steps=None):
    """car/YOLO.py:209-240 -> (sum(area), A, 4) fp32 [left, top, right, bottom]."""
    steps = steps or decode.init_steps(spec)
    area = decode.init_area(spec, steps)
    H, W = spec["size"]
    out = []
    for i, anchors in enumerate(spec["all_anchors"]):
        anchors = np.asarray(anchors, np.float32)
        n, a, step = len(anchors), area[i], f32(steps[i])
        h, w = anchors[:, 0], anchors[:, 1]
        x_num, y_num = int(W / step), int(H / step)
        # nd.arange(start, 1, step, repeat): start + i*step evaluated in fp32
        ys = (f32(step / f32(H) / f32(2.)) + np.arange(y_num, dtype=np.float32) * f32(step / f32(H))).astype(np.float32)
        xs = (f32(step / f32(W) / f32(2.)) + np.arange(x_num, dtype=np.float32) * f32(step / f32(W))).astype(np.float32)
        y = np.repeat(ys, n * x_num)
        hh = np.tile(h, a)
        top = (y - f32(0.5) * hh).reshape(a, n, 1)
        bot = (y + f32(0.5) * hh).reshape(a, n, 1)
        x = np.repeat(xs, n)
        ww = np.tile(w, x_num)
        left = np.tile(x - f32(0.5) * ww, y_num).reshape(a, n, 1)
        right = np.tile(x + f32(0.5) * ww, y_num).reshape(a, n, 1)
        out.append(np.concatenate([left, top, right, bot], axis=-1).astype(np.float32))
    return np.concatenate(out, axis=0)


def get_iou_mode2(ltrb, L):
    """yolo_gluon.get_iou mode 2, fp32 op by op."""
    l, t, r, b = (ltrb[..., i] for i in range(4))
    l2 = f32(L[2] - L[4] / f32(2)); t2 = f32(L[1] - L[3] / f32(2))
    r2 = f32(L[2] + L[4] / f32(2)); b2 = f32(L[1] + L[3] / f32(2))
    iw = np.maximum(np.minimum(r2, r) - np.maximum(l2, l), f32(0)).astype(np.float32)
    ih = np.maximum(np.minimum(b2, b) - np.maximum(t2, t), f32(0)).astype(np.float32)
    inters = (iw * ih).astype(np.float32)
    pa = ((r - l) * (b - t)).astype(np.float32)
    ta = f32(L[3] * L[4])
    with np.errstate(invalid="ignore", divide="ignore"):
        return (inters / ((pa + ta).astype(np.float32) - inters)).astype(np.float32)


def find_best(spec, ltrb, L, steps=None):
    """car/YOLO.py:401-448 -> (flat index = pixel*A + anchor, [ty, tx, th, tw] fp32)."""
    steps = steps or decode.init_steps(spec)
    area = decode.init_area(spec, steps)
    A = len(spec["all_anchors"][0])
    H, W = spec["size"]
    L = np.asarray(L, np.float32)
    best = int(np.argmax(get_iou_mode2(ltrb, L).reshape(-1)))
    px, anc = best // A, best % A
    if px >= sum(area):
        px = sum(area) - 1
    bl = ltrb[px, anc]
    a0, layer = 0, 0
    for i, a in enumerate(area):
        a0 += a
        if px < a0:
            layer = i
            break
    step = f32(steps[layer])
    by = f32(L[1] - f32(f32(bl[3] + bl[1]) / f32(2)))
    sy = np.clip(f32(f32(f32(by * f32(H)) / step) + f32(0.5)), f32(0.0001), f32(0.9999)).astype(np.float32)
    bx = f32(L[2] - f32(f32(bl[2] + bl[0]) / f32(2)))
    sx = np.clip(f32(f32(f32(bx * f32(W)) / step) + f32(0.5)), f32(0.0001), f32(0.9999)).astype(np.float32)
    inv = lambda v: f32(-np.log(f32(f32(1) / v - f32(1)), dtype=np.float32))
    anchors = np.asarray(spec["all_anchors"], np.float32)
    th = f32(np.log(f32(L[3] / anchors[layer, anc, 0]), dtype=np.float32))
    tw = f32(np.log(f32(L[4] / anchors[layer, anc, 1]), dtype=np.float32))
    return px * A + anc, np.asarray([inv(sy), inv(sx), th, tw], np.float32)


def loss_mask(spec, labels, steps=None):
    """car/YOLO.py:450-480 -> ([score, yx, hw, rotate, class] dense targets, mask, assignment (B,obj) flat idx or -1)."""
    labels = np.asarray(labels, np.float32)
    B, nobj = labels.shape[:2]
    area = decode.init_area(spec, steps)
    a, n = sum(area), len(spec["all_anchors"][0])
    nc = labels.shape[2] - 6
    ltrb = default_ltrb(spec, steps)
    mask = np.zeros((B, a, n, 1), np.float32)
    score = np.zeros((B, a, n, 1), np.float32)
    yx = np.zeros((B, a, n, 2), np.float32)
    hw = np.zeros((B, a, n, 2), np.float32)
    rot = np.zeros((B, a, n, 1), np.float32)
    cls = np.zeros((B, a, n, nc), np.float32)
    assign = np.full((B, nobj), -1, np.int32)
    for b in range(B):
        for j, L in enumerate(labels[b]):
            if L[0] < 0:
                continue
            flat, box = find_best(spec, ltrb, L, steps)
            px, anc = flat // n, flat % n
            mask[b, px, anc] = 1.0
            score[b, px, anc] = 1.0
            yx[b, px, anc] = box[:2]
            hw[b, px, anc] = box[2:]
            rot[b, px, anc] = L[5]
            cls[b, px, anc] = L[6:]
            assign[b, j] = flat
    return [score, yx, hw, rot, cls], mask, assign


def _mean_nb(t):
    return t.reshape(t.shape[0], -1).mean(dim=1)


def get_loss(spec, heads, targets, mask, hp, car_rotate=False):
    """car/YOLO.py:491-498 with the gluon loss definitions; torch fp32 so autograd gives the reference gradients.
    heads: list of (B,HW_s,A,C) torch tensors (may require grad).  Returns the 5 losses, each (B,)."""
    x = torch.cat(list(heads), dim=1)
    sp = spec["slice_point"]
    xs, i = [], 0
    for pt in sp:
        xs.append(x[..., i:pt]); i = pt
    y = [torch.as_tensor(t) for t in targets]
    m = torch.as_tensor(mask)
    sw = torch.where(m > 0, torch.full_like(m, hp["positive_weight"]), torch.full_like(m, hp["negative_weight"]))

    def logistic(p, lab, w):
        lab = 2 * lab - 1
        l = torch.relu(-p * lab) + torch.nn.functional.softplus(-torch.abs(p * lab))
        return _mean_nb(l * w)

    def huber(p, lab, w, rho=1.0):
        d = torch.abs(lab - p)
        l = torch.where(d > rho, d - 0.5 * rho, (0.5 / rho) * d * d)
        return _mean_nb(l * w)

    def softmax_ce(p, lab, w):
        lp = torch.log_softmax(p, dim=-1)
        l = -(lp * lab).sum(dim=-1, keepdim=True)
        return _mean_nb(l * w)

    sc = hp["scale"]
    rot_lr = sc["rotate"] if car_rotate else 0.0
    return (logistic(xs[0], y[0], sw * sc["score"]), huber(xs[1], y[1], m * sc["box_yx"]), huber(xs[2], y[2], m * sc["box_hw"]),
            huber(xs[3], y[3], m * rot_lr), softmax_ce(xs[4], y[4], m * sc["class"]))


V1_HPARAMS = dict(scale={"score": 0.1, "box_yx": 0.01, "box_hw": 10.0, "rotate": 0.0, "class": 0.3},
                  positive_weight=1.0, negative_weight=0.1)        # car/v1/spec.yaml:27-33


def synthetic_labels(batch, num_class, nobj=1, seed=99, p_box=0.5):
    """with p=0.5 a box [cls, y~U(.2,.8), x~U(.2,.8), h~U(.15,.7), w~U(.15,.7), 0, soft class dist], else all -1
    (render_car.py:80,123-132)."""
    rng = np.random.default_rng(seed)
    lab = np.full((batch, nobj, 6 + num_class), -1.0, np.float32)
    for b in range(batch):
        for j in range(nobj):
            if rng.random() < p_box:
                c = int(rng.integers(0, num_class))
                d = np.exp(-((np.arange(num_class) - c) ** 2) / 2.0)
                lab[b, j, :6] = [c, rng.uniform(.2, .8), rng.uniform(.2, .8), rng.uniform(.15, .7), rng.uniform(.15, .7), 0.0]
                lab[b, j, 6:] = (d / d.sum()).astype(np.float32)
    return lab


# ---------------------------------------------------------------------------------------------------------
# one training step (car/YOLO.py:350-399 + gluon.Trainer/Adam), torch autograd as the reference
# ---------------------------------------------------------------------------------------------------------
TRAINABLE = ("weight", "gamma", "beta", "bias")


def train_step(net, spec, params, x, labels, hp, lr=0.001, batch_size=None, adam=None, t=1, car_rotate=False,
               beta1=0.9, beta2=0.999, eps=1e-8, dtype=torch.float32, lp_labels=None, lp_hp=None):
    """Forward in train mode (batch-statistics BN, per device), targets, the five losses, ``sum(losses).backward()``,
    then ``trainer.step(batch_size)``: grad * (1/batch_size) -> MXNet ``adam_update`` with the bias-corrected lr
    (``lr * sqrt(1-beta2^t)/(1-beta1^t)``, epsilon outside the square root, wd 0).  Single device (the multi-context sum of
    the reference is the all-reduce of the B200 build).  Returns dict(losses, grads, params, adam)."""
    from . import nets
    import math
    batch_size = batch_size or x.shape[0]
    tp = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=k.rsplit(".", 1)[1] in TRAINABLE) for k, v in params.items()}
    out, new_stats = nets.forward(net, spec, tp, torch.as_tensor(x).to(dtype), train=True)          # dtype=float64: the noise reference of the parity tests
    heads = out if net == "carnet" else out[0]
    targets, mask, assign = loss_mask(spec, labels)
    losses = get_loss(spec, heads, targets, mask, hp, car_rotate)
    if lp_labels is not None:                                  # car_and_LP/YOLO.py:265-304: the LP losses join the backward
        lp_x = out[1][0]
        lt, lm = loss_mask_LP(spec, lp_labels, spec["size"][0] // lp_x.shape[1])
        losses = tuple(losses) + tuple(get_loss_LP(spec, lp_x, lt, lm, lp_hp))
    sum(l.sum() for l in losses).backward()
    grads = {k: v.grad.numpy().copy() for k, v in tp.items() if v.requires_grad}
    adam = adam or {k: (np.zeros_like(g), np.zeros_like(g)) for k, g in grads.items()}
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    new_params, new_adam = {}, {}
    for k, v in params.items():
        if k in grads:
            g = grads[k] * np.float32(1.0 / batch_size)
            m, vv = adam[k]
            m = (beta1 * m + (1 - beta1) * g).astype(np.float32)
            vv = (beta2 * vv + (1 - beta2) * g * g).astype(np.float32)
            new_params[k] = (np.asarray(v, np.float32) - np.float32(lr_t) * m / (np.sqrt(vv) + np.float32(eps))).astype(np.float32)
            new_adam[k] = (m, vv)
        elif k in new_stats:
            new_params[k] = new_stats[k].detach().numpy().astype(np.float32)
        else:
            new_params[k] = np.asarray(v, np.float32)
    return dict(losses=np.stack([l.detach().numpy() for l in losses]), grads=grads, params=new_params, adam=new_adam, assign=assign,
                heads=[h.detach().numpy() for h in heads])


# ---------------------------------------------------------------------------------------------------------
# licence-plate pose targets and losses (licence_plate/LP_detection.py:259-360, car_and_LP/YOLO.py:124-131,265-304)
# ---------------------------------------------------------------------------------------------------------
LP_V1_HPARAMS = dict(scale={"LP_score": 0.1, "LP_xy": 10.0, "LP_z": 1.0, "LP_r": 0.1, "LP_class": 0.0},
                     LP_positive_weight=1.0, LP_negative_weight=0.1)          # car_and_LP/v1/spec.yaml:43,51-52


def find_best_LP(spec, L, step):
    """LP_detection.py:259-283: cell (h, w) = clip(int(L[8] / step)), clip(int(L[7] / step)); targets X/1000, Y/1000, Z/1000 and
    sigma^-1(r_i / r_max_i / 2 + 0.5) (nd_inv_sigmoid = -log(1/x - 1), yolo_gluon.py:365-367), fp32 op by op."""
    f = np.float32
    h_max, w_max = spec["size"][0] // step - 1, spec["size"][1] // step - 1
    hf = int(np.clip(int(f(L[8]) / f(step)), 0, h_max))
    wf = int(np.clip(int(f(L[7]) / f(step)), 0, w_max))
    t = [f(L[1]) / f(1000.0), f(L[2]) / f(1000.0), f(L[3]) / f(1000.0)]
    for i in range(3):
        rmax = f(f(spec["LP_r_max"][i]) * f(np.pi) / f(180.0))
        x = f(f(f(L[4 + i]) / rmax) / f(2.0)) + f(0.5)
        t.append(f(-np.log(f(f(1.0) / x) - f(1.0), dtype=np.float32)))
    return (hf, wf), np.asarray(t, np.float32)


def loss_mask_LP(spec, labels, step):
    """LP_detection.py:285-313 -> ([score, pose_xy, pose_z, pose_r, LP_class] dense targets (B,Hs,Ws,k), mask).
    A later label overwrites pose/score of an earlier one in the same cell; class one-hots accumulate (the reference never clears them)."""
    labels = np.asarray(labels, np.float32)
    B = labels.shape[0]
    hs, ws, nc = spec["size"][0] // step, spec["size"][1] // step, spec["LP_num_class"]
    score, mask = np.zeros((B, hs, ws, 1), np.float32), np.zeros((B, hs, ws, 1), np.float32)
    xy, z, r = np.zeros((B, hs, ws, 2), np.float32), np.zeros((B, hs, ws, 1), np.float32), np.zeros((B, hs, ws, 3), np.float32)
    cls = np.zeros((B, hs, ws, nc), np.float32)
    for b in range(B):
        for L in labels[b]:
            if L[0] < 0:
                continue
            (hf, wf), p = find_best_LP(spec, L, step)
            score[b, hf, wf] = 1.0
            mask[b, hf, wf] = 1.0
            xy[b, hf, wf] = p[:2]
            z[b, hf, wf] = p[2]
            r[b, hf, wf] = p[3:]
            cls[b, hf, wf, int(L[-1])] = 1
    return [score, xy, z, r, cls], mask


def get_loss_LP(spec, lp_x, targets, mask, hp):
    """LP_detection.py:354-360 with the gluon loss definitions; lp_x: (B,Hs,Ws,ch) torch tensor (may require grad)."""
    sp = spec["LP_slice_point"]
    xs, i = [], 0
    for pt in sp:
        xs.append(lp_x[..., i:pt]); i = pt
    y = [torch.as_tensor(t) for t in targets]
    m = torch.as_tensor(mask)
    sw = torch.where(m > 0, torch.full_like(m, hp["LP_positive_weight"]), torch.full_like(m, hp["LP_negative_weight"]))
    sc = hp["scale"]

    def logistic(p, lab, w):
        lab = 2 * lab - 1
        return _mean_nb((torch.relu(-p * lab) + torch.nn.functional.softplus(-torch.abs(p * lab))) * w)

    def huber(p, lab, w, rho=1.0):
        d = torch.abs(lab - p)
        return _mean_nb(torch.where(d > rho, d - 0.5 * rho, (0.5 / rho) * d * d) * w)

    def softmax_ce(p, lab, w):
        return _mean_nb(-(torch.log_softmax(p, dim=-1) * lab).sum(dim=-1, keepdim=True) * w)

    return (logistic(xs[0], y[0], sw * sc["LP_score"]), huber(xs[1], y[1], m * sc["LP_xy"]), huber(xs[2], y[2], m * sc["LP_z"]),
            huber(xs[3], y[3], m * sc["LP_r"]), softmax_ce(xs[4], y[4], m * sc["LP_class"]))


def synthetic_LP_labels(batch, size, nobj=1, seed=5, p_box=0.7, num_class=3):
    """LP label rows [flag, X, Y, Z (mm), r1, r2, r3 (rad), pixel x, pixel y, class] (render layout read by _find_best_LP)."""
    rng = np.random.default_rng(seed)
    lab = np.full((batch, nobj, 10), -1.0, np.float32)
    for b in range(batch):
        for j in range(nobj):
            if rng.random() < p_box:
                lab[b, j] = [1, rng.uniform(-2000, 2000), rng.uniform(-1000, 1000), rng.uniform(3000, 20000), rng.uniform(-0.6, 0.6),
                             rng.uniform(-0.8, 0.8), rng.uniform(-0.6, 0.6), rng.uniform(0, size[1] - 1), rng.uniform(0, size[0] - 1),
                             rng.integers(0, num_class)]
    return lab
