"""Oracle restatement (numpy fp32) of the reference's anchor decode + box selection.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

Follows (file:line under /root/reference):
  * ``car/YOLO.py:112-155``  _init_step / _init_area / _init_syxhw   (per-box s,y,x,h,w tables)
  * ``car/YOLO.py:841-849``  merge_and_slice                          (concat scales, slice channels)
  * ``car/YOLO.py:552-566``  _yxhw_to_ltrb
  * ``car/YOLO.py:568-597``  predict                                  (per-image top-1 argmax)
  * ``car_and_LP/YOLO.py:133-169`` predict_LP / LP_pose_activation    (argmax of sigmoid(score))
  * ``licence_plate/LP_detection.py:147-162`` predict_LP              (argmax of the RAW score, image 0)
  * ``yolo_modules/yolo_gluon.py:127-168`` get_iou                    (IoU form used by the NMS extension)

Sigmoid: MXNet's CPU kernel is ``1.0f / (1.0f + expf(-x))`` (mshadow_op::sigmoid).  ``expf`` is restated
as the correctly-rounded fp32 exponential (exp in float64, rounded once to fp32), which is what glibc's
``expf`` returns in all but astronomically rare cases; the add and the divide are IEEE fp32.  The CUDA
kernel computes the identical expression so that saturation ties (x >~ 17 -> exactly 1.0f) and the
first-occurrence rule of ``argmax`` select the same flat index.

NMS is NOT in the reference; ``nms`` below is the north-star extension defined in
SURVEY.md 8(c): greedy, score-descending, class-aware, IoU from ``get_iou``'s ltrb form, ties broken by
the lower flat index.  Its first kept box always equals ``predict``'s top-1.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def sigmoid32(x):
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(over="ignore"):
        e = np.exp(-x.astype(np.float64)).astype(np.float32)   # correctly rounded expf(-x); +inf on overflow like expf
    return (f32(1.0) / (f32(1.0) + e)).astype(np.float32)


def exp32(x):
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(over="ignore"):
        return np.exp(x.astype(np.float64)).astype(np.float32)


def init_steps(spec):
    """car/YOLO.py:112-116."""
    num_downsample = len(spec["layers"])
    npyr = len(spec["all_anchors"])
    start = num_downsample - npyr + 1
    return [2 ** (start + i) for i in range(npyr)]


def init_area(spec, steps=None):
    """car/YOLO.py:118-121 (int() truncation preserved)."""
    steps = steps or init_steps(spec)
    h, w = spec["size"]
    return [int(h * w / step ** 2) for step in steps]


def init_syxhw(spec, steps=None):
    """car/YOLO.py:123-155 -> five (1, sum(area), A, 1) fp32 tables."""
    steps = steps or init_steps(spec)
    area = init_area(spec, steps)
    size = spec["size"]
    n = len(spec["all_anchors"][0])
    tot = sum(area)
    S, Y, X, Hh, Ww = (np.zeros((1, tot, n, 1), np.float32) for _ in range(5))
    a0 = 0
    for i, anchors in enumerate(spec["all_anchors"]):
        a, step = area[i], steps[i]
        x_num = int(size[1] / step)
        y_num = int(size[0] / step)
        s = np.full(a * n, step, np.float32)
        y = np.repeat(np.arange(0, size[0], step, dtype=np.float32), n * x_num)
        x = np.tile(np.repeat(np.arange(0, size[1], step, dtype=np.float32), n), y_num)
        hw = np.tile(np.asarray(anchors, np.float32), (a, 1))
        h, w = hw[:, 0], hw[:, 1]
        S[0, a0:a0 + a] = s.reshape(a, n, 1)
        Y[0, a0:a0 + a] = y.reshape(a, n, 1)
        X[0, a0:a0 + a] = x.reshape(a, n, 1)
        Hh[0, a0:a0 + a] = h.reshape(a, n, 1)
        Ww[0, a0:a0 + a] = w.reshape(a, n, 1)
        a0 += a
    return S, Y, X, Hh, Ww


def merge_and_slice(all_output, points):
    """car/YOLO.py:841-849."""
    out = np.concatenate(all_output, axis=1)
    x, i = [], 0
    for pt in points:
        x.append(out[..., i:pt])
        i = pt
    return x


def yxhw_to_ltrb(spec, yxhw, tables):
    """car/YOLO.py:552-566 (all ops in fp32, same operation order)."""
    S, Y, X, Hh, Ww = tables
    ty, tx, th, tw = (yxhw[..., i:i + 1] for i in range(4))
    by = ((sigmoid32(ty) * S + Y) / f32(spec["size"][0])).astype(np.float32)
    bx = ((sigmoid32(tx) * S + X) / f32(spec["size"][1])).astype(np.float32)
    bh = (exp32(th) * Hh).astype(np.float32)
    bw = (exp32(tw) * Ww).astype(np.float32)
    bh2 = bh / f32(2)
    bw2 = bw / f32(2)
    l, r, t, b = bx - bw2, bx + bw2, by - bh2, by + bh2
    return np.concatenate([l, t, r, b], axis=-1).astype(np.float32)


def decode_all(spec, heads, steps=None):
    """Everything ``predict`` builds before the per-image loop: (B, boxes, 6+num_class) rows
    [sigmoid(score), l, t, r, b, rotate_raw, class_logits_raw...] plus the score tensor."""
    tables = init_syxhw(spec, steps)
    heads = [np.asarray(h, np.float32) for h in heads]
    sl = merge_and_slice(heads, spec["slice_point"])
    score = sigmoid32(sl[0])
    box = yxhw_to_ltrb(spec, np.concatenate([sl[1], sl[2]], axis=-1), tables)
    out = np.concatenate([score, box, sl[3], sl[4]], axis=-1).astype(np.float32)
    B = out.shape[0]
    return out.reshape(B, -1, out.shape[-1]), score.reshape(B, -1)


def predict(spec, heads, steps=None, return_index=False):
    """car/YOLO.py:568-597 -> (B, 6+num_class) fp32 [score, y, x, h, w, rotate, class...]."""
    rows, score = decode_all(spec, heads, steps)
    preds, idxs = [], []
    for i in range(rows.shape[0]):
        best = int(np.argmax(score[i]))                 # first occurrence wins (MXNet argmax)
        pred = rows[i, best].copy()
        y = (pred[2] + pred[4]) / f32(2)
        x = (pred[1] + pred[3]) / f32(2)
        h = pred[4] - pred[2]
        w = pred[3] - pred[1]
        pred[1:5] = [y, x, h, w]
        preds.append(pred)
        idxs.append(best)
    preds = np.stack(preds).astype(np.float32)
    if return_index:
        return preds, np.asarray(idxs, np.int32)
    return preds


def lp_pose_activation(spec, v):
    """car_and_LP/YOLO.py:159-169."""
    out = np.zeros(6, np.float32)
    out[0:3] = v[0:3] * f32(1000)
    for i in range(3):
        d = (sigmoid32(v[i + 3]) - f32(0.5)) * f32(2) * f32(spec["LP_r_max"][i])
        out[i + 3] = f32(d) * f32(math.pi) / f32(180.0)
    return out


def predict_LP_batch(spec, lp_out, return_index=False):
    """car_and_LP/YOLO.py:133-157: lp_out (B, Hs, Ws, 10) -> (B, 7)."""
    lp_out = np.asarray(lp_out, np.float32)
    B = lp_out.shape[0]
    flat = lp_out.reshape(B, -1, lp_out.shape[-1])
    score = sigmoid32(flat[..., 0])
    preds, idxs = [], []
    for i in range(B):
        best = int(np.argmax(score[i]))
        pred = np.concatenate([[score[i, best]], flat[i, best, 1:7]]).astype(np.float32)
        pred[1:7] = lp_pose_activation(spec, pred[1:7])
        preds.append(pred)
        idxs.append(best)
    preds = np.stack(preds).astype(np.float32)
    if return_index:
        return preds, np.asarray(idxs, np.int32)
    return preds


def predict_LP_single(spec, net_out, return_index=False):
    """licence_plate/LP_detection.py:147-162: NCHW (B,10,H,W) -> (10,) for image 0; argmax on RAW score."""
    out = np.asarray(net_out, np.float32).transpose(0, 2, 3, 1)[0]
    best = int(np.argmax(out[:, :, 0].reshape(-1)))
    pred = out.reshape(-1, spec["LP_slice_point"][-1])[best].copy()
    pred[0] = sigmoid32(pred[0])
    pred[1:4] *= f32(1000)
    for i in range(3):
        p = (sigmoid32(pred[i + 4]) - f32(0.5)) * f32(2) * f32(spec["LP_r_max"][i])
        pred[i + 4] = f32(p) * f32(math.pi) / f32(180.0)
    if return_index:
        return pred, best
    return pred


def iou_ltrb(a, b):
    """yolo_gluon.py:158-167 in ltrb form, fp32."""
    il, it = np.maximum(a[0], b[0]), np.maximum(a[1], b[1])
    ir, ib = np.minimum(a[2], b[2]), np.minimum(a[3], b[3])
    iw = np.maximum(f32(ir - il), f32(0))
    ih = np.maximum(f32(ib - it), f32(0))
    inter = f32(iw * ih)
    area_a = f32(f32(a[2] - a[0]) * f32(a[3] - a[1]))
    area_b = f32(f32(b[2] - b[0]) * f32(b[3] - b[1]))
    return f32(inter / f32(f32(area_a + area_b) - inter))


def nms(spec, heads, score_thr=0.5, iou_thr=0.45, max_out=100, max_cand=1024, steps=None):
    """North-star extension (not in the reference).  Per image: candidates = boxes with
    sigmoid(score) > score_thr, ordered by (score desc, flat index asc), truncated to ``max_cand``;
    greedy suppression of later same-class boxes with IoU > iou_thr; at most ``max_out`` kept.
    If nothing passes the threshold the top-1 box is kept alone (so kept[0] == predict's index always).
    Returns list per image of (rows (K, 6+num_class) in predict's row format, indices (K,) int32)."""
    rows, score = decode_all(spec, heads, steps)
    res = []
    for i in range(rows.shape[0]):
        sc = score[i]
        cand = np.nonzero(sc > f32(score_thr))[0]
        if cand.size == 0:
            cand = np.asarray([int(np.argmax(sc))])
        order = np.lexsort((cand, -sc[cand].astype(np.float64)))
        cand = cand[order][:max_cand]
        boxes = rows[i, cand, 1:5]
        cls = np.argmax(rows[i, cand, 6:], axis=-1) if rows.shape[-1] > 6 else np.zeros(len(cand), int)
        alive = np.ones(len(cand), bool)
        keep = []
        for a in range(len(cand)):
            if not alive[a]:
                continue
            keep.append(a)
            if len(keep) >= max_out:
                break
            for b in range(a + 1, len(cand)):
                if alive[b] and cls[b] == cls[a] and iou_ltrb(boxes[a], boxes[b]) > f32(iou_thr):
                    alive[b] = False
        out = rows[i, cand[keep]].copy()
        l, t, r, b = (out[:, k].copy() for k in (1, 2, 3, 4))
        out[:, 1] = (t + b) / f32(2)
        out[:, 2] = (l + r) / f32(2)
        out[:, 3] = b - t
        out[:, 4] = r - l
        res.append((out.astype(np.float32), cand[keep].astype(np.int32)))
    return res
