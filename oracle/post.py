"""Oracle restatement (numpy / math, double precision like the reference) of the post-decode consumers.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

Follows (file:line under /root/reference):
  * ``car/video_node.py:36-38,244-252`` and ``yolo_modules/yolo_cv.py:85-94``  cls2ang: softmax -> circular mean angle
  * ``yolo_modules/licence_plate_render/__init__.py:340-377``                  ProjectRectangle6D.__call__ / projection_matrix
  * ``:379-402`` add_edges: cv2.getPerspectiveTransform + cv2.warpPerspective (cv2 itself is the reference library here)
"""
import math

import numpy as np


def cls2ang(confidence, logits):
    step = 360 // len(logits)
    cos_offset = np.array([math.cos(x * math.pi / 180) for x in range(0, 360, step)])
    sin_offset = np.array([math.sin(x * math.pi / 180) for x in range(0, 360, step)])
    x = np.asarray(logits, np.float32)
    prob = np.exp(x) / np.sum(np.exp(x), axis=0)
    c = sum(cos_offset * prob)
    s = sum(sin_offset * prob)
    return math.atan2(s, c), confidence * (s ** 2 + c ** 2) ** 0.5


def project_rectangle(pose, fx, fy, cx, cy):
    X, Y, Z, r1, r2, r3 = [float(v) for v in pose[:6]]
    sin, cos = math.sin, math.cos
    a = sin(r1) * cos(r2) * 84.0
    b = sin(r1) * sin(r2) * cos(r3) * 84.0
    c = sin(r2) * 199.5
    d = sin(r3) * cos(r1) * 84.0
    e = cos(r2) * cos(r3) * 199.5
    f = sin(r1) * sin(r2) * sin(r3) * 84.0
    g = sin(r3) * cos(r2) * 199.5
    h = cos(r1) * cos(r3) * 84.0
    ans = np.array([
        [cx * (Z + a - c) + fx * (X + b - d + e), cx * (Z + a + c) + fx * (X + b - d - e), cx * (Z - a + c) + fx * (X - b + d - e),
         cx * (Z - a - c) + fx * (X - b + d + e)],
        [cy * (Z + a - c) + fy * (Y + f + g + h), cy * (Z + a + c) + fy * (Y + f - g + h), cy * (Z - a + c) + fy * (Y - f - g - h),
         cy * (Z - a - c) + fy * (Y - f + g - h)],
        [Z + a - c, Z + a + c, Z - a + c, Z - a - c]])
    pts = np.zeros((4, 2))
    for i in range(4):
        pts[i, 0] = ans[0, i] / ans[2, i]
        pts[i, 1] = ans[1, i] / ans[2, i]
    return pts.astype(np.float32)


def add_edges(img, corner_pts, LP_size=(160, 380)):
    import cv2
    LP_corner = np.float32([[LP_size[1], LP_size[0]], [0, LP_size[0]], [0, 0], [LP_size[1], 0]])
    M = cv2.getPerspectiveTransform(np.asarray(corner_pts, np.float32), LP_corner)
    return cv2.warpPerspective(img, M, (LP_size[1], LP_size[0]))
