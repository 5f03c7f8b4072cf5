"""yolo_b200 - B200-native (sm_100a) YOLOv3 detection hot path behind the Python surface of n8886919/YOLO.

Hand-written CUDA behind a C ABI (``include/yolo_b200.h``, ``libyolo_b200.so``); Python is the host side
the reference's drivers expect (``YOLO.predict``, ``predict_LP``, ``net.forward``).  No CPU fallback.
"""
from ._lib import YoloError, load as load_library  # noqa: F401
from .api import (Net, NDArray, Trainer, decode_top1, decode_nms, decode_lp, loss_targets, lp_loss_targets, azimuth, lp_corners, resize_u8,  # noqa: F401
                  lp_unwarp)
from .drivers import YOLO, YOLO_dense, CarLPYOLO, LicencePlateDetectioin, ProjectRectangle6D, cls2ang, get_ctx, init_NN  # noqa: F401
from . import mxnet_io  # noqa: F401

__all__ = ["Net", "NDArray", "Trainer", "decode_top1", "decode_nms", "decode_lp", "loss_targets", "lp_loss_targets", "azimuth", "resize_u8", "lp_corners", "lp_unwarp", "YOLO", "YOLO_dense", "ProjectRectangle6D", "cls2ang", "CarLPYOLO", "LicencePlateDetectioin",
           "get_ctx", "init_NN", "mxnet_io", "YoloError", "load_library"]
