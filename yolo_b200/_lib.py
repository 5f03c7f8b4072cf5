"""ctypes binding of ``libyolo_b200.so`` (C ABI declared in ``include/yolo_b200.h``).

The library is the product: there is NO CPU or PyTorch fallback.  If the shared object is missing
or cannot be loaded this module raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YOLO_B200_LIB") or os.path.join(_HERE, "libyolo_b200.so")      # override: A/B runs of two builds

MAX_STAGES, MAX_SCALES, MAX_ANCHORS, MAX_BLOCKS = 8, 3, 8, 8
NET_CARNET, NET_CARLPNET, NET_LPDENSENET, NET_DEBUGCONV, NET_CARDENSENET = 0, 1, 2, 3, 4
PREC_FP32, PREC_BF16, PREC_BF16X6, PREC_FP16X3 = 0, 1, 2, 3
IN_NCHW_F32, IN_NHWC_U8 = 0, 1
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x6": PREC_BF16X6, "fp16x3": PREC_FP16X3}


class YoloSpec(C.Structure):
    _fields_ = [
        ("net_type", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("n_layers", C.c_int32), ("layers", C.c_int32 * MAX_STAGES), ("channels", C.c_int32 * (MAX_STAGES + 1)),
        ("n_scales", C.c_int32), ("n_anchors", C.c_int32),
        ("anchors", ((C.c_float * 2) * MAX_ANCHORS) * MAX_SCALES),
        ("channels_per_anchor", C.c_int32), ("lp_channels", C.c_int32), ("lp_r_max", C.c_float * 3),
        ("lp_num_class", C.c_int32),
        ("num_init_features", C.c_int32), ("growth_rate", C.c_int32), ("bn_size", C.c_int32),
        ("n_blocks", C.c_int32), ("block_config", C.c_int32 * MAX_BLOCKS),
        ("precision", C.c_int32), ("max_batch", C.c_int32),
    ]


class DecodeGeom(C.Structure):
    _fields_ = [
        ("height", C.c_int32), ("width", C.c_int32),
        ("n_scales", C.c_int32), ("n_anchors", C.c_int32), ("channels_per_anchor", C.c_int32),
        ("step", C.c_int32 * MAX_SCALES),
        ("anchors", ((C.c_float * 2) * MAX_ANCHORS) * MAX_SCALES),
    ]


class LossParams(C.Structure):
    _fields_ = [("scale_score", C.c_float), ("scale_box_yx", C.c_float), ("scale_box_hw", C.c_float), ("scale_rotate", C.c_float),
                ("scale_class", C.c_float), ("positive_weight", C.c_float), ("negative_weight", C.c_float), ("car_rotate", C.c_int32)]


class LpLossParams(C.Structure):
    _fields_ = [("scale_score", C.c_float), ("scale_xy", C.c_float), ("scale_z", C.c_float), ("scale_r", C.c_float), ("scale_class", C.c_float),
                ("positive_weight", C.c_float), ("negative_weight", C.c_float)]


class NmsParams(C.Structure):
    _fields_ = [("score_thr", C.c_float), ("iou_thr", C.c_float), ("max_out", C.c_int32), ("max_cand", C.c_int32)]


# every symbol include/yolo_b200.h declares: name -> (restype, argtypes)
_VP, _I, _SZ = C.c_void_p, C.c_int, C.c_size_t
SYMBOLS = {
    "yolo_version": (C.c_char_p, []),
    "yolo_create": (_I, [C.POINTER(YoloSpec), _I, C.POINTER(_VP)]),
    "yolo_destroy": (_I, [_VP]),
    "yolo_last_error": (C.c_char_p, [_VP]),
    "yolo_param_count": (_I, [_VP]),
    "yolo_param_info": (_I, [_VP, _I, C.POINTER(C.c_char_p), C.POINTER(C.c_int32 * 4), C.POINTER(C.c_int32)]),
    "yolo_load_param": (_I, [_VP, C.c_char_p, _VP, _SZ]),
    "yolo_finalize_params": (_I, [_VP, _VP]),
    "yolo_workspace_bytes": (_SZ, [_VP, _I]),
    "yolo_set_workspace": (_I, [_VP, _VP, _SZ]),
    "yolo_output_count": (_I, [_VP]),
    "yolo_output_shape": (_I, [_VP, _I, C.POINTER(C.c_int32 * 4), C.POINTER(C.c_int32)]),
    "yolo_forward": (_I, [_VP, _VP, _I, _I, C.POINTER(_VP), _VP]),
    "yolo_check_saturation": (_I, [_VP, C.POINTER(C.c_int32), _VP]),
    "yolo_debug_activation": (_I, [_VP, C.c_char_p, _I, _VP, _SZ]),
    "yolo_debug_wgrad": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _VP, _I, _VP]),
    "yolo_decode_top1": (_I, [C.POINTER(DecodeGeom), C.POINTER(_VP), _I, _VP, _VP, _VP]),
    "yolo_decode_nms": (_I, [C.POINTER(DecodeGeom), C.POINTER(_VP), _I, C.POINTER(NmsParams), _VP, _VP, _VP, _VP]),
    "yolo_decode_lp": (_I, [_VP, _I, _I, _I, _I, _I, C.POINTER(C.c_float * 3), _VP, _VP, _VP]),
    "yolo_loss_scratch_bytes": (_SZ, [_I, _I]),
    "yolo_loss_targets": (_I, [C.POINTER(DecodeGeom), C.POINTER(_VP), _VP, _I, _I, C.POINTER(LossParams), _VP, _VP, C.POINTER(_VP), _VP, _VP]),
    "yolo_resize_u8": (_I, [_VP, _I, _I, _I, _VP, _I, _I, _VP]),
    "yolo_azimuth": (_I, [_VP, _I, _I, _I, _VP, _VP, _VP]),
    "yolo_lp_corners": (_I, [_VP, _I, _I, _I, C.POINTER(C.c_double * 4), C.c_float, C.c_float, _VP, _VP]),
    "yolo_lp_unwarp": (_I, [_VP, _I, _I, _I, _I, _VP, _I, _I, _VP, _VP, _VP]),
    "yolo_lp_loss_targets": (_I, [_VP, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_float * 3), _VP, _I, _I, C.POINTER(LpLossParams), _VP, _VP, _VP]),
    "yolo_train_forward_backward_lp": (_I, [_VP, _VP, _I, _VP, _I, _I, C.POINTER(LossParams), _VP, _I, _I, C.POINTER(LpLossParams), _VP, _VP]),
    "yolo_train_flat_size": (_SZ, [_VP]),
    "yolo_train_init": (_I, [_VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "yolo_train_set_bn_momentum": (_I, [_VP, C.c_float]),
    "yolo_nccl_unique_id": (_I, [_VP]),
    "yolo_train_comm_init": (_I, [_VP, _VP, _I, _I, _SZ]),
    "yolo_train_forward_backward": (_I, [_VP, _VP, _I, _VP, _I, _I, C.POINTER(LossParams), _VP, _VP]),
    "yolo_train_apply": (_I, [_VP, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _VP]),
    "yolo_get_param": (_I, [_VP, C.c_char_p, _VP, _SZ, _I]),
    "yolo_predict_host": (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP]),
    "yolo_last_launch_count": (_I, [_VP]),
    "yolo_conv_flops_per_image": (C.c_double, [_VP]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it is absent - build with ``python -m yolo_b200.build``."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run `python -m yolo_b200.build`). "
            "yolo_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class YoloError(RuntimeError):
    pass


def check(rc, handle=None):
    if rc == 0:
        return
    msg = load().yolo_last_error(handle)
    raise YoloError(f"yolo_b200 error {rc}: {msg.decode() if msg else '?'}")
