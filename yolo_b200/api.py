"""Python host side of yolo_b200: thin objects over the C ABI.

PyTorch is used only as the device-memory container (workspace, inputs, outputs) and for streams.
``Net`` mirrors the executor the reference drivers hold in ``self.net`` (``yolo_gluon.init_executor``,
yolo_modules/yolo_gluon.py:204-242): ``net.forward(is_train=False, data=...)`` returns the list of head
tensors, each with ``.shape`` / ``.wait_to_read()`` / ``.asnumpy()`` / ``.copy()`` like an mx NDArray.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (DecodeGeom, LossParams, LpLossParams, NmsParams, YoloSpec, check, IN_NCHW_F32, IN_NHWC_U8, NET_CARNET, NET_CARLPNET,
                   NET_LPDENSENET, NET_DEBUGCONV, NET_CARDENSENET, PRECISIONS)

NET_TYPES = {"carnet": NET_CARNET, "carlpnet": NET_CARLPNET, "lpdensenet": NET_LPDENSENET, "debugconv": NET_DEBUGCONV,
             "cardensenet": NET_CARDENSENET}


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("yolo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class NDArray:
    """Minimal mx.nd.NDArray look-alike over a torch CUDA tensor."""

    def __init__(self, t: torch.Tensor):
        self.t = t

    @property
    def shape(self):
        return tuple(self.t.shape)

    def wait_to_read(self):
        torch.cuda.current_stream(self.t.device).synchronize()
        return self

    def asnumpy(self):
        return self.t.detach().cpu().numpy()

    def copy(self):
        return NDArray(self.t.clone())

    def __len__(self):
        return self.t.shape[0]


def _as_tensor(x):
    return x.t if isinstance(x, NDArray) else x


def init_steps(spec):
    """car/YOLO.py:112-116 (DenseNet variants: the single head sits at stride 2^(len(block_config) + 1))."""
    if "layers" not in spec:
        return [2 ** (len(spec["block_config"]) + 1)]
    nd, npyr = len(spec["layers"]), len(spec["all_anchors"])
    start = nd - npyr + 1
    return [2 ** (start + i) for i in range(npyr)]


def make_c_spec(net_type, spec, precision="fp32", max_batch=1):
    s = YoloSpec()
    s.net_type = NET_TYPES[net_type]
    s.height, s.width = int(spec["size"][0]), int(spec["size"][1])
    s.precision = PRECISIONS[precision]
    s.max_batch = int(max_batch)
    s.bn_size = int(spec.get("bn_size", 4))
    if net_type in ("carnet", "carlpnet"):
        layers, channels, anchors = spec["layers"], spec["channels"], spec["all_anchors"]
        if len(layers) != len(channels) - 1:
            raise ValueError("len(channels) should equal to len(layers) + 1, given {} vs {}".format(len(channels), len(layers)))
        if len(layers) > _lib.MAX_STAGES or len(anchors) > _lib.MAX_SCALES or len(anchors[0]) > _lib.MAX_ANCHORS:
            raise ValueError("spec exceeds compiled limits")
        sp = list(spec["slice_point"])
        if sp[:4] != [1, 3, 5, 6]:
            raise ValueError(f"slice_point {sp} unsupported: the decode kernel is built for [1,3,5,6,C] (car/v1/spec.yaml:6)")
        s.n_layers = len(layers)
        for i, v in enumerate(layers):
            s.layers[i] = int(v)
        for i, v in enumerate(channels):
            s.channels[i] = int(v)
        s.n_scales, s.n_anchors = len(anchors), len(anchors[0])
        for i, sc in enumerate(anchors):
            for j, (ah, aw) in enumerate(sc):
                s.anchors[i][j][0], s.anchors[i][j][1] = float(ah), float(aw)
        s.channels_per_anchor = int(sp[-1])
    if net_type == "cardensenet":          # car/YOLO.py:864-893: one scale, A anchors, C channels per anchor
        anchors = spec["all_anchors"]
        s.n_scales, s.n_anchors = 1, len(anchors[0])
        for j, (ah, aw) in enumerate(anchors[0]):
            s.anchors[0][j][0], s.anchors[0][j][1] = float(ah), float(aw)
        s.channels_per_anchor = int(spec["slice_point"][-1])
    if net_type in ("carlpnet", "lpdensenet"):
        s.lp_channels = int(spec["LP_slice_point"][-1])
        for i in range(3):
            s.lp_r_max[i] = float(spec["LP_r_max"][i])
        s.lp_num_class = int(spec["LP_num_class"])
    if net_type == "debugconv":          # kernel unit-test harness (see include/yolo_b200.h)
        s.channels[0], s.channels[1] = int(spec["cin"]), int(spec["cout"])
        for i, key in enumerate(("k", "stride", "pad", "act", "residual", "bn")):
            s.layers[i] = int(spec[key])
    if net_type in ("lpdensenet", "cardensenet"):
        s.num_init_features, s.growth_rate = int(spec["num_init_features"]), int(spec["growth_rate"])
        cfg = spec["block_config"]
        s.n_blocks = len(cfg)
        for i, v in enumerate(cfg):
            s.block_config[i] = int(v)
    return s


def make_geom(spec, steps=None):
    g = DecodeGeom()
    g.height, g.width = int(spec["size"][0]), int(spec["size"][1])
    anchors = spec["all_anchors"]
    g.n_scales, g.n_anchors = len(anchors), len(anchors[0])
    g.channels_per_anchor = int(spec["slice_point"][-1])
    steps = steps or init_steps(spec)
    for i, st in enumerate(steps):
        g.step[i] = int(st)
    for i, sc in enumerate(anchors):
        for j, (ah, aw) in enumerate(sc):
            g.anchors[i][j][0], g.anchors[i][j][1] = float(ah), float(aw)
    return g


class Net:
    """A network instance on one GPU (one process per GPU; see yolo_b200.parallel for sharding)."""

    def __init__(self, net_type, spec, precision="fp32", max_batch=1, device=0):
        _require_cuda()
        self.lib = _lib.load()
        self.net_type, self.spec, self.precision, self.max_batch = net_type, spec, precision, int(max_batch)
        self.device = torch.device("cuda", device)
        self._h = C.c_void_p()
        cs = make_c_spec(net_type, spec, precision, max_batch)
        check(self.lib.yolo_create(C.byref(cs), device, C.byref(self._h)))
        self._ws = None
        self._pinned = None
        self.out_shapes = []
        for i in range(self.lib.yolo_output_count(self._h)):
            shp, nd = (C.c_int32 * 4)(), C.c_int32()
            check(self.lib.yolo_output_shape(self._h, i, C.byref(shp), C.byref(nd)), self._h)
            self.out_shapes.append(tuple(shp[k] for k in range(nd.value)))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self.lib.yolo_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ---- parameters (yolo_gluon.init_NN, yolo_modules/yolo_gluon.py:172-201) -------------------------
    def param_shapes(self):
        out = []
        for i in range(self.lib.yolo_param_count(self._h)):
            name, shp, nd = C.c_char_p(), (C.c_int32 * 4)(), C.c_int32()
            check(self.lib.yolo_param_info(self._h, i, C.byref(name), C.byref(shp), C.byref(nd)), self._h)
            out.append((name.value.decode(), tuple(shp[k] for k in range(nd.value))))
        return out

    def load_params(self, params: dict):
        for name, shape in self.param_shapes():
            if name not in params:
                raise KeyError(f"missing parameter {name}")
            a = np.ascontiguousarray(np.asarray(params[name], dtype=np.float32))
            if tuple(a.shape) != shape:
                raise ValueError(f"{name}: expected shape {shape}, got {tuple(a.shape)}")
            check(self.lib.yolo_load_param(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.size), self._h)
        with torch.cuda.device(self.device):
            check(self.lib.yolo_finalize_params(self._h, _stream_ptr(self.device)), self._h)
            self._ensure_workspace()
        return self

    def _ensure_workspace(self):
        if self._ws is None:
            n = self.lib.yolo_workspace_bytes(self._h, self.max_batch)
            # zero-filled: tiles of the tensor-core kernels may run past the last image of a smaller batch and must read finite values
            self._ws = torch.zeros(max(n, 1024) + 1024, dtype=torch.uint8, device=self.device)
            ptr = (self._ws.data_ptr() + 1023) & ~1023
            check(self.lib.yolo_set_workspace(self._h, C.c_void_p(ptr), n), self._h)

    # ---- forward (car/video_node.py:230-231) ------------------------------------------------------
    def _to_device(self, data):
        data = _as_tensor(data)
        if isinstance(data, np.ndarray):
            # two pinned staging buffers, each guarded by an event recorded after its H2D copy: the host never overwrites
            # a buffer whose asynchronous copy may still be in flight (pipelined inference / training loops)
            t = torch.from_numpy(data)
            nbytes = t.numel() * t.element_size()
            if self._pinned is None:
                self._pinned, self._pin_turn = [None, None], 0
            i = self._pin_turn
            self._pin_turn ^= 1
            slot = self._pinned[i]
            if slot is None or slot[0].numel() < nbytes:
                slot = self._pinned[i] = [torch.empty(nbytes, dtype=torch.uint8).pin_memory(), None]
            if slot[1] is not None:
                slot[1].synchronize()
            stage = slot[0][:nbytes].view(t.dtype).view(t.shape)
            stage.copy_(t)
            with torch.cuda.device(self.device):
                d = stage.to(self.device, non_blocking=True)
                slot[1] = torch.cuda.Event()
                slot[1].record(torch.cuda.current_stream(self.device))
            return d
        if not data.is_cuda:
            return data.to(self.device, non_blocking=True)
        return data

    def forward(self, is_train=False, data=None):
        if is_train:
            # the reference records the forward for autograd (car/YOLO.py:383-384) and calls backward on the losses; here the whole
            # step is one fused call because BatchNorm statistics, losses and gradients never leave the device
            raise NotImplementedError("train-mode forward is part of the fused step: use YOLO._train_batch / Trainer.forward_backward")
        x = self._to_device(data)
        H, W = self.spec["size"]
        if x.dtype == torch.float32 and x.dim() == 4 and tuple(x.shape[1:]) == (3, H, W):
            layout = IN_NCHW_F32
        elif x.dtype == torch.uint8 and x.dim() == 4 and tuple(x.shape[1:]) == (H, W, 3):
            layout = IN_NHWC_U8
        else:
            raise ValueError(f"input must be (B,3,{H},{W}) float32 or (B,{H},{W},3) uint8, got {tuple(x.shape)} {x.dtype}")
        x = x.contiguous()
        B = x.shape[0]
        with torch.cuda.device(self.device):
            outs = [torch.empty((B,) + s, dtype=torch.float32, device=self.device) for s in self.out_shapes]
            ptrs = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
            check(self.lib.yolo_forward(self._h, C.c_void_p(x.data_ptr()), B, layout, ptrs, _stream_ptr(self.device)), self._h)
        self._keep = x                       # keep the input alive until the stream has consumed it
        return [NDArray(o) for o in outs]

    def predict_host(self, frames: torch.Tensor):
        """One C call: H2D + forward + decode_top1 + D2H (yolo_predict_host).  ``frames`` is a (pinned) host tensor."""
        B = frames.shape[0]
        layout = IN_NCHW_F32 if frames.dtype == torch.float32 else IN_NHWC_U8
        C_ = int(self.spec["slice_point"][-1])
        rows = np.empty((B, C_), np.float32)
        idx = np.empty((B,), np.int32)
        with torch.cuda.device(self.device):
            check(self.lib.yolo_predict_host(self._h, C.c_void_p(frames.data_ptr()), B, layout, rows.ctypes.data_as(C.c_void_p),
                                             idx.ctypes.data_as(C.c_void_p), _stream_ptr(self.device)), self._h)
        return rows, idx

    def saturated(self):
        """fp16x3 range flags since the last call (bit 0: an activation exceeded 65504, bit 1: a re-packed weight did); clears them."""
        v = C.c_int32(0)
        with torch.cuda.device(self.device):
            check(self.lib.yolo_check_saturation(self._h, C.byref(v), _stream_ptr(self.device)), self._h)
        return int(v.value)

    def activation(self, name, shape):
        """Copy an internal activation (oracle layer name) to host as NCHW fp32 of the given shape (parity tests)."""
        out = np.empty(shape, np.float32)
        check(self.lib.yolo_debug_activation(self._h, name.encode(), shape[0], out.ctypes.data_as(C.c_void_p), out.size), self._h)
        return out

    @property
    def launches(self):
        return self.lib.yolo_last_launch_count(self._h)

    @property
    def conv_flops_per_image(self):
        return self.lib.yolo_conv_flops_per_image(self._h)


# ---- decode entry points (no network handle needed) ------------------------------------------------------
def _head_ptrs(heads):
    ts = [_as_tensor(h) for h in heads]
    for t in ts:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("heads must be contiguous float32 CUDA tensors")
    return ts, (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def decode_top1(spec, heads, steps=None, out=None):
    """car/YOLO.py:568-597 as one kernel.  Returns (rows (B,C) cuda fp32, idx (B,) cuda int32); ``out=(rows, idx)`` reuses buffers."""
    lib = _lib.load()
    ts, ptrs = _head_ptrs(heads)
    B, dev = ts[0].shape[0], ts[0].device
    g = make_geom(spec, steps)
    if out is not None:
        rows, idx = out
    else:
        rows = torch.empty((B, g.channels_per_anchor), dtype=torch.float32, device=dev)
        idx = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.yolo_decode_top1(C.byref(g), ptrs, B, C.c_void_p(rows.data_ptr()), C.c_void_p(idx.data_ptr()), _stream_ptr(dev)))
    return rows, idx


def decode_nms(spec, heads, score_thr=0.5, iou_thr=0.45, max_out=100, max_cand=1024, steps=None, out=None):
    """Fused decode + class-aware NMS.  Returns (rows (B,max_out,C), idx (B,max_out), count (B,)); entries beyond count[b] are 0 / -1
    unless ``out=(rows, idx, count)`` passes caller-owned buffers (then only the first count[b] entries are defined)."""
    lib = _lib.load()
    ts, ptrs = _head_ptrs(heads)
    B, dev = ts[0].shape[0], ts[0].device
    g = make_geom(spec, steps)
    p = NmsParams(float(score_thr), float(iou_thr), int(max_out), int(max_cand))
    if out is not None:
        rows, idx, cnt = out
    else:
        rows = torch.zeros((B, max_out, g.channels_per_anchor), dtype=torch.float32, device=dev)
        idx = torch.full((B, max_out), -1, dtype=torch.int32, device=dev)
        cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.yolo_decode_nms(C.byref(g), ptrs, B, C.byref(p), C.c_void_p(rows.data_ptr()), C.c_void_p(idx.data_ptr()),
                                  C.c_void_p(cnt.data_ptr()), _stream_ptr(dev)))
    return rows, idx, cnt


def decode_lp(lp, mode, r_max):
    """mode 0: (B,Hs,Ws,ch) NHWC, sigmoid-score argmax -> (B,7); mode 1: (B,ch,Hs,Ws) NCHW, raw-score argmax -> (B,ch)."""
    lib = _lib.load()
    t = _as_tensor(lp)
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 4):
        raise ValueError("lp must be a contiguous 4-d float32 CUDA tensor")
    if mode == 0:
        B, hs, ws, ch = t.shape
        nout = 7
    else:
        B, ch, hs, ws = t.shape
        nout = ch
    rm = (C.c_float * 3)(*[float(v) for v in r_max])
    rows = torch.empty((B, nout), dtype=torch.float32, device=t.device)
    idx = torch.empty((B,), dtype=torch.int32, device=t.device)
    with torch.cuda.device(t.device):
        check(lib.yolo_decode_lp(C.c_void_p(t.data_ptr()), B, hs, ws, ch, mode, C.byref(rm), C.c_void_p(rows.data_ptr()),
                                 C.c_void_p(idx.data_ptr()), _stream_ptr(t.device)))
    return rows, idx


def resize_u8(frames, size):
    """``cv2.resize(frame, (W, H))`` (bilinear, OpenCV's fixed-point arithmetic) on the GPU: frames (B,h,w,3) / (h,w,3) uint8 (numpy or
    cuda tensor), size = (H, W) like spec['size'] -> cuda uint8 (B,H,W,3), ready for ``Net.forward`` (car/video_node.py:150)."""
    lib = _lib.load()
    t = _as_tensor(frames)
    t = torch.as_tensor(t)
    if t.dim() == 3:
        t = t[None]
    if not t.is_cuda:
        t = t.cuda()
    t = t.contiguous()
    if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
        raise ValueError("frames must be uint8 (B,h,w,3)")
    out = torch.empty((t.shape[0], int(size[0]), int(size[1]), 3), dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        check(lib.yolo_resize_u8(C.c_void_p(t.data_ptr()), t.shape[0], t.shape[1], t.shape[2], C.c_void_p(out.data_ptr()), out.shape[1], out.shape[2],
                                 _stream_ptr(t.device)))
    return out


def azimuth(rows, n_class=24):
    """car/video_node.py:244-252 / yolo_cv.cls2ang: rows (B,C) cuda fp32 from decode_top1 -> (angle (B,), radius (B,)) cuda fp32."""
    lib = _lib.load()
    t = _as_tensor(rows)
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 2):
        raise ValueError("rows must be a contiguous (B, C) float32 CUDA tensor")
    ang = torch.empty((t.shape[0],), dtype=torch.float32, device=t.device)
    rad = torch.empty_like(ang)
    with torch.cuda.device(t.device):
        check(lib.yolo_azimuth(C.c_void_p(t.data_ptr()), t.shape[0], t.shape[1], int(n_class), C.c_void_p(ang.data_ptr()), C.c_void_p(rad.data_ptr()),
                               _stream_ptr(t.device)))
    return ang, rad


def lp_corners(poses, intrinsics, x_scale=1.0, y_scale=1.0, pose_offset=1):
    """ProjectRectangle6D.__call__ (licence_plate_render/__init__.py:340-377): poses (B,>=offset+6) cuda fp32 rows holding
    [X, Y, Z (mm), r1, r2, r3 (rad)] at ``pose_offset`` (1 for predict_LP rows); intrinsics = (fx, fy, cx, cy) -> (B,4,2) pixel corners."""
    lib = _lib.load()
    t = _as_tensor(poses)
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 2):
        raise ValueError("poses must be a contiguous (B, n) float32 CUDA tensor")
    out = torch.empty((t.shape[0], 4, 2), dtype=torch.float32, device=t.device)
    k = (C.c_double * 4)(*[float(v) for v in intrinsics])
    with torch.cuda.device(t.device):
        check(lib.yolo_lp_corners(C.c_void_p(t.data_ptr()), t.shape[0], t.shape[1], int(pose_offset), C.byref(k), float(x_scale), float(y_scale),
                                  C.c_void_p(out.data_ptr()), _stream_ptr(t.device)))
    return out


def lp_unwarp(img, corners, LP_size=(160, 380)):
    """``add_edges`` (licence_plate_render/__init__.py:379-402): img (H,W,3) or (B,H,W,3) cuda uint8, corners (B,4,2) cuda fp32 ->
    (clipped plates (B, LP_size[0], LP_size[1], 3) uint8, ok (B,) int32)."""
    lib = _lib.load()
    im, cn = _as_tensor(img), _as_tensor(corners)
    if not (im.is_cuda and im.dtype == torch.uint8 and im.is_contiguous() and im.dim() in (3, 4) and im.shape[-1] == 3):
        raise ValueError("img must be a contiguous uint8 CUDA tensor (H,W,3) or (B,H,W,3)")
    if not (cn.is_cuda and cn.dtype == torch.float32 and cn.is_contiguous() and tuple(cn.shape[1:]) == (4, 2)):
        raise ValueError("corners must be a contiguous (B,4,2) float32 CUDA tensor")
    B = cn.shape[0]
    batched = im.dim() == 4
    if batched and im.shape[0] != B:
        raise ValueError("one frame per plate, or a single frame")
    H, W = im.shape[-3], im.shape[-2]
    out = torch.empty((B, LP_size[0], LP_size[1], 3), dtype=torch.uint8, device=im.device)
    ok = torch.empty((B,), dtype=torch.int32, device=im.device)
    with torch.cuda.device(im.device):
        check(lib.yolo_lp_unwarp(C.c_void_p(im.data_ptr()), B, int(batched), H, W, C.c_void_p(cn.data_ptr()), LP_size[0], LP_size[1],
                                 C.c_void_p(out.data_ptr()), C.c_void_p(ok.data_ptr()), _stream_ptr(im.device)))
    return out, ok


def loss_targets(spec, heads, labels, scale, positive_weight, negative_weight, car_rotate=False, with_grad=False, steps=None):
    """Targets + losses of car/YOLO.py:385-394 (``_loss_mask`` + ``_get_loss``) on the GPU.

    heads: list of (B,HW_s,A,C) cuda fp32; labels: (B,n_obj,6+num_class) cuda/np fp32.
    Returns (losses (5,B) cuda fp32 in spec `loss_name` order, assignment (B,n_obj) cuda int32, dheads or None)."""
    lib = _lib.load()
    ts, ptrs = _head_ptrs(heads)
    B, dev = ts[0].shape[0], ts[0].device
    lab = torch.as_tensor(labels, dtype=torch.float32).to(dev).contiguous()
    if lab.dim() != 3 or lab.shape[0] != B or lab.shape[2] != int(spec["slice_point"][-1]):
        raise ValueError(f"labels must be (B, n_obj, 6+num_class), got {tuple(lab.shape)}")
    n_obj = lab.shape[1]
    g = make_geom(spec, steps)
    p = LossParams(float(scale["score"]), float(scale["box_yx"]), float(scale["box_hw"]), float(scale["rotate"]), float(scale["class"]),
                   float(positive_weight), float(negative_weight), int(bool(car_rotate)))
    scratch = torch.empty(lib.yolo_loss_scratch_bytes(B, n_obj), dtype=torch.uint8, device=dev)
    losses = torch.empty((5, B), dtype=torch.float32, device=dev)
    assign = torch.empty((B, n_obj), dtype=torch.int32, device=dev)
    dheads, dptrs = None, None
    if with_grad:
        dheads = [torch.empty_like(t) for t in ts]
        dptrs = (C.c_void_p * len(dheads))(*[t.data_ptr() for t in dheads])
    with torch.cuda.device(dev):
        check(lib.yolo_loss_targets(C.byref(g), ptrs, C.c_void_p(lab.data_ptr()), B, n_obj, C.byref(p), C.c_void_p(scratch.data_ptr()),
                                    C.c_void_p(losses.data_ptr()), dptrs, C.c_void_p(assign.data_ptr()), _stream_ptr(dev)))
    return losses, assign, dheads


def _lp_params(scale, positive_weight, negative_weight):
    return LpLossParams(float(scale["LP_score"]), float(scale["LP_xy"]), float(scale["LP_z"]), float(scale["LP_r"]), float(scale["LP_class"]),
                        float(positive_weight), float(negative_weight))


def lp_loss_targets(lp_map, labels, step, r_max, scale, positive_weight, negative_weight, nchw=False, with_grad=False):
    """`_loss_mask_LP` + `_get_loss_LP` (licence_plate/LP_detection.py:285-313,354-360) on the GPU.
    lp_map: (B,Hs,Ws,ch) cuda fp32 (``nchw=False``, CarLPNet) or (B,ch,Hs,Ws) (LPDenseNet output); labels: (B,n_obj,>=10).
    Returns (losses (5,B) cuda fp32 in order LP_score, LP_xy, LP_z, LP_r, LP_class, d(sum)/d(lp_map) or None)."""
    lib = _lib.load()
    t = _as_tensor(lp_map)
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 4):
        raise ValueError("lp_map must be a contiguous 4-d float32 CUDA tensor")
    B = t.shape[0]
    ch, hs, ws = (t.shape[1], t.shape[2], t.shape[3]) if nchw else (t.shape[3], t.shape[1], t.shape[2])
    lab = torch.as_tensor(labels, dtype=torch.float32).to(t.device).contiguous()
    if lab.dim() != 3 or lab.shape[0] != B or lab.shape[2] < 10:
        raise ValueError(f"LP labels must be (B, n_obj, >= 10), got {tuple(lab.shape)}")
    rm = (C.c_float * 3)(*[float(v) for v in r_max])
    p = _lp_params(scale, positive_weight, negative_weight)
    losses = torch.empty((5, B), dtype=torch.float32, device=t.device)
    dlp = torch.empty_like(t) if with_grad else None
    with torch.cuda.device(t.device):
        check(lib.yolo_lp_loss_targets(C.c_void_p(t.data_ptr()), int(bool(nchw)), B, hs, ws, ch, int(step), C.byref(rm), C.c_void_p(lab.data_ptr()),
                                       lab.shape[1], lab.shape[2], C.byref(p), C.c_void_p(losses.data_ptr()),
                                       C.c_void_p(dlp.data_ptr()) if with_grad else None, _stream_ptr(t.device)))
    return losses, dlp


class Trainer:
    """Data-parallel training of a ``Net`` (CARNET / CARLPNET, precision fp16x3 = fp32-grade on the tensor cores), one process per
    GPU - the B200 shape of ``_init_train`` + ``_train_batch`` + ``gluon.Trainer(..., 'adam').step(batch_size)``
    (car/YOLO.py:157-207, 350-399).

    The flat parameter / gradient / Adam buffers are torch tensors (device-memory containers).  When torch.distributed is
    initialised with more than one rank, the library joins its own NCCL communicator (the 128-byte id travels through
    torch.distributed) and sums the gradient over ranks bucket by bucket while the backward runs - the kvstore reduction inside
    ``trainer.step``.  BatchNorm statistics stay per GPU like in the reference."""

    def __init__(self, net, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, bucket_bytes=64 << 20, join_nccl=True):
        self.net, self.lib = net, net.lib
        self.lr, self.beta1, self.beta2, self.eps = float(learning_rate), float(beta1), float(beta2), float(epsilon)
        n = self.lib.yolo_train_flat_size(net._h)
        dev = net.device
        self.P, self.G, self.M, self.V = (torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(4))
        with torch.cuda.device(dev):
            check(self.lib.yolo_train_init(net._h, C.c_void_p(self.P.data_ptr()), C.c_void_p(self.G.data_ptr()), C.c_void_p(self.M.data_ptr()),
                                           C.c_void_p(self.V.data_ptr()), n, _stream_ptr(dev)), net._h)
        net._trainer = self              # the handle reads the flat buffers: keep them alive as long as the net
        self.world, self.rank, self.nccl_joined = 1, 0, False
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.world, self.rank = dist.get_world_size(), dist.get_rank()
            if join_nccl:
                self._join_nccl(dist, int(bucket_bytes))

    def _join_nccl(self, dist, bucket_bytes):
        dev = self.net.device
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            check(self.lib.yolo_nccl_unique_id(C.c_void_p(uid.data_ptr())))
        on_gpu = dist.get_backend() == "nccl"
        t = uid.to(dev) if on_gpu else uid
        dist.broadcast(t, src=0)
        uid = t.cpu().contiguous()
        with torch.cuda.device(dev):
            check(self.lib.yolo_train_comm_init(self.net._h, C.c_void_p(uid.data_ptr()), self.rank, self.world, bucket_bytes), self.net._h)
        self.nccl_joined = True

    def set_bn_momentum(self, momentum):
        check(self.lib.yolo_train_set_bn_momentum(self.net._h, float(momentum)), self.net._h)

    def forward_backward(self, images, labels, scale, positive_weight, negative_weight, car_rotate=False, lp_labels=None,
                         lp_positive_weight=1.0, lp_negative_weight=0.1):
        """images: (b,3,H,W) fp32 or (b,H,W,3) uint8; labels: (b,n_obj,6+num_class).  Returns the (5,b) losses (cuda); with
        ``lp_labels`` (b,n_lp,>=10) on a CarLPNet the five LP losses join the backward and (10,b) is returned (car rows first).
        With the library's communicator joined the gradient buffer holds the sum over ranks when this stream reaches it."""
        net = self.net
        x = net._to_device(images).contiguous()
        layout = IN_NCHW_F32 if x.dtype == torch.float32 else IN_NHWC_U8
        lab = torch.as_tensor(labels, dtype=torch.float32).to(net.device).contiguous()
        B, n_obj = x.shape[0], lab.shape[1]
        p = LossParams(float(scale["score"]), float(scale["box_yx"]), float(scale["box_hw"]), float(scale["rotate"]), float(scale["class"]),
                       float(positive_weight), float(negative_weight), int(bool(car_rotate)))
        with torch.cuda.device(net.device):
            if lp_labels is None:
                losses = torch.empty((5, B), dtype=torch.float32, device=net.device)
                check(self.lib.yolo_train_forward_backward(net._h, C.c_void_p(x.data_ptr()), layout, C.c_void_p(lab.data_ptr()), B, n_obj, C.byref(p),
                                                           C.c_void_p(losses.data_ptr()), _stream_ptr(net.device)), net._h)
                self._keep = (x, lab)
            else:
                lpl = torch.as_tensor(lp_labels, dtype=torch.float32).to(net.device).contiguous()
                lpp = _lp_params(scale, lp_positive_weight, lp_negative_weight)
                losses = torch.empty((10, B), dtype=torch.float32, device=net.device)
                check(self.lib.yolo_train_forward_backward_lp(net._h, C.c_void_p(x.data_ptr()), layout, C.c_void_p(lab.data_ptr()), B, n_obj, C.byref(p),
                                                              C.c_void_p(lpl.data_ptr()), lpl.shape[1], lpl.shape[2], C.byref(lpp),
                                                              C.c_void_p(losses.data_ptr()), _stream_ptr(net.device)), net._h)
                self._keep = (x, lab, lpl)
        return losses

    def allreduce_grads(self):
        """Sum of the gradients over ranks.  A no-op when the library reduced them during the backward (the default)."""
        if self.nccl_joined or self.world == 1:
            return
        import torch.distributed as dist
        dist.all_reduce(self.G, op=dist.ReduceOp.SUM)

    def step(self, batch_size):
        """``trainer.step(batch_size)``: gradients (already summed over ranks) are rescaled by 1/batch_size, then Adam."""
        with torch.cuda.device(self.net.device):
            check(self.lib.yolo_train_apply(self.net._h, self.lr, self.beta1, self.beta2, self.eps, 1.0 / float(batch_size),
                                            _stream_ptr(self.net.device)), self.net._h)

    def get_param(self, name, shape, grad=False):
        out = np.empty(shape, np.float32)
        check(self.lib.yolo_get_param(self.net._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.size, int(grad)), self.net._h)
        return out

    def calibrate_bn(self, images):
        """Set every BatchNorm's running statistics to the batch statistics of ``images`` (one train-mode forward with momentum 0;
        parameters untouched) - what synthetic random weights need to behave like a trained net."""
        net = self.net
        C_ = int(net.spec["slice_point"][-1])
        lab = np.full((images.shape[0], 1, C_), -1.0, np.float32)
        sc = {"score": 0.0, "box_yx": 0.0, "box_hw": 0.0, "rotate": 0.0, "class": 0.0}
        self.set_bn_momentum(0.0)
        self.forward_backward(images, lab, sc, 1.0, 1.0)
        self.set_bn_momentum(0.9)
        lr = self.lr
        self.lr = 0.0
        self.step(1)                     # lr 0: parameters stay, the inference epilogues are refolded with the new statistics
        self.lr = lr
        torch.cuda.current_stream(net.device).synchronize()
