"""Host-side mirrors of the reference's task drivers for the hot path (same names, arguments and
return values, so callers such as car/video_node.py keep working):

  * ``YOLO``                    car/YOLO.py:48-110, predict :568-597
  * ``CarLPYOLO``               car_and_LP/YOLO.py:98-169 (class ``YOLO`` there), predict_LP :133-157
  * ``LicencePlateDetectioin``  licence_plate/LP_detection.py:100-162 (spelling is the reference's)

Only the hot path lives here (construction from spec.yaml, ``net.forward``, ``predict`` / ``predict_LP``);
rendering, ROS, tensorboard and checkpoint rotation are out of scope.
There is no CPU fallback: unlike ``yolo_gluon.get_ctx`` (yolo_modules/yolo_gluon.py:393-396) a missing GPU
raises.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np
import torch
import yaml

from . import api


def get_ctx(gpu):
    """yolo_gluon.get_ctx (yolo_modules/yolo_gluon.py:380-408): list of device indices; no CPU fallback."""
    if not torch.cuda.is_available():
        raise RuntimeError("NO GPU be Detected! yolo_b200 has no CPU fallback")
    if isinstance(gpu, str):
        gpu = [g for g in gpu.replace(" ", "").split(",") if g != ""]
    ctx = [int(g) for g in gpu if int(g) < torch.cuda.device_count()]
    if not ctx:
        raise RuntimeError(f"GPU index error: {gpu}")
    return ctx


def init_NN(target, weight, ctx=None, seed=0):
    """yolo_gluon.init_NN (yolo_modules/yolo_gluon.py:172-201): load an MXNet ``collect_params().save`` / export ``.params`` file into
    ``target`` (an ``api.Net``); if that fails, fall back to Xavier initialisation like the reference does (and say so)."""
    from . import mxnet_io, synth
    print("use pretrain weight: %s" % weight)
    try:
        params = mxnet_io.load_gluon_params(weight, target.net_type, target.spec, target.param_shapes())
        target.load_params(params)
        print("Load Pretrain Successfully")
        return True
    except Exception as e:      # noqa: BLE001  (the reference catches everything here)
        print("Load Pretrain Failed, Use Xavier initializer")
        print(str(e).split("\n")[0])
        C = target.spec["slice_point"][-1] if "slice_point" in target.spec else None
        target.load_params(synth.random_params(target.param_shapes(), seed=seed, channels_per_anchor=C))
        return False


def _load_spec(args, spec):
    if spec is None:
        with open(os.path.join(args.version, "spec.yaml")) as f:
            spec = yaml.safe_load(f)
    return spec


def _args(args, **kw):
    if args is None:
        args = SimpleNamespace(version=None, mode="video", gpu="0", weight=None, record=0, trt=0)
    for k, v in kw.items():
        if not hasattr(args, k):
            setattr(args, k, v)
    return args


class YOLO:
    """Vehicle box + azimuth detector (car/YOLO.py).  ``params`` is a dict name -> fp32 array in the
    canonical naming of the C library (``net.param_shapes()``); without it the net must be loaded later
    through ``self.net.load_params``."""

    net_type = "carnet"

    def __init__(self, args=None, spec=None, params=None, precision="fp32", max_batch=1, gpu=None):
        args = _args(args, gpu="0", mode="video")
        self.ctx = get_ctx(args.gpu if gpu is None else str(gpu))
        spec = _load_spec(args, spec)
        for key in spec:                                   # car/YOLO.py:58-59
            setattr(self, key, spec[key])
        self.spec = spec
        self.all_anchors = np.asarray(self.all_anchors, np.float32)
        self.num_class = len(self.classes) if hasattr(self, "classes") else self.slice_point[-1] - 6
        if self.slice_point[-1] != 6 + self.num_class:
            raise ValueError("slice_point[-1] must equal 6 + len(classes)")
        self._init_step()
        self._init_area()
        self.version = getattr(args, "version", None)
        self.precision, self.max_batch = precision, max_batch
        self._init_net(spec, params, getattr(args, "weight", None))

    def _init_step(self):                                  # car/YOLO.py:112-116
        self.steps = api.init_steps(self.spec)

    def _init_area(self):                                  # car/YOLO.py:118-121
        h, w = self.size
        self.area = [int(h * w / step ** 2) for step in self.steps]

    def _init_net(self, spec, params, weight=None):        # car/YOLO.py:91-110
        self.net = api.Net(self.net_type, spec, self.precision, self.max_batch, self.ctx[0])
        if params is not None:
            self.net.load_params(params)
        elif weight:                                       # args.weight: an MXNet .params file (car/YOLO.py:98-100)
            init_NN(self.net, weight, self.ctx)

    # -------------------- Validation Part -------------------- #
    def predict(self, batch_out, mode="top1", score_thr=0.5, iou_thr=0.45, max_out=100, max_cand=1024, return_index=False):
        """car/YOLO.py:568-597: list of (B,HW_s,A,C) heads -> np.float32 (B, 6+num_class)
        rows [score, y, x, h, w, rotate, class logits...].  ``mode='nms'`` is the north-star extension and
        returns (rows (B,max_out,C), indices (B,max_out), counts (B,)) as numpy arrays."""
        heads = list(batch_out)[: len(self.steps)]
        if mode == "top1":
            rows, idx = api.decode_top1(self.spec, heads, self.steps)
            if return_index:
                both = torch.cat([rows, idx.view(torch.float32).view(-1, 1)], dim=1)   # one D2H like asnumpy(); bit-cast
                both = both.cpu().numpy()
                return both[:, :-1].copy(), np.ascontiguousarray(both[:, -1]).view(np.int32)
            return rows.cpu().numpy()
        if mode == "nms":
            rows, idx, cnt = api.decode_nms(self.spec, heads, score_thr, iou_thr, max_out, max_cand, self.steps)
            return rows.cpu().numpy(), idx.cpu().numpy(), cnt.cpu().numpy()
        raise ValueError("mode must be 'top1' or 'nms'")


    # -------------------- Training Part -------------------- #
    def _loss_mask_and_get_loss(self, y_, car_by, car_rotate=False, with_grad=False):
        """GPU replacement of ``_loss_mask`` + ``_score_weight`` + ``_get_loss`` (car/YOLO.py:385-392, 450-498) for one device:
        y_ = list of head tensors from ``net.forward``, car_by = labels (b, num_object, 6+num_class).
        Returns the five per-image losses in ``loss_name`` order as a (5, b) CUDA tensor (+ head gradients if asked)."""
        losses, assign, dheads = api.loss_targets(self.spec, list(y_)[: len(self.steps)], car_by, self.scale, self.positive_weight,
                                                  self.negative_weight, car_rotate, with_grad, self.steps)
        return (losses, assign, dheads) if with_grad else (losses, assign)


    def _init_train(self):
        """car/YOLO.py:157-207: Adam(learning_rate from the spec); batch_size is the GLOBAL batch (batch_size *= len(ctx))."""
        import torch.distributed as dist
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.global_batch_size = int(getattr(self, "batch_size", self.max_batch)) * world
        self.trainer = api.Trainer(self.net, learning_rate=getattr(self, "learning_rate", 0.001))
        self.backward_counter = getattr(self, "train_counter_start", 0)

    def _train_batch(self, bxs, car_bys, car_rotate=False):
        """car/YOLO.py:350-399.  ``bxs`` / ``car_bys`` are lists with ONE entry (this process's GPU; the reference passes one
        entry per context).  Side effects like the reference: parameters updated, ``backward_counter`` advanced;
        returns None.  The per-image losses of the last step are kept in ``self.last_losses`` ((5,b), loss_name order)."""
        if not hasattr(self, "trainer"):
            self._init_train()
        if len(bxs) != 1 or len(car_bys) != 1:
            raise ValueError(f"one process per GPU (torchrun contract, INTEGRATION.md): pass this rank's slice as single-element lists, "
                             f"got {len(bxs)} entries for ctx {self.ctx}")
        self.last_losses = self.trainer.forward_backward(bxs[0], car_bys[0], self.scale, self.positive_weight, self.negative_weight, car_rotate)
        self.trainer.allreduce_grads()
        self.trainer.step(self.global_batch_size)
        self.backward_counter += 1

    train_step = _train_batch          # north-star name of the same call


class YOLO_dense(YOLO):
    """car/YOLO.py:864-937: the DenseNet-backed detector (``CarDenseNet``, car/utils.py:48-61) - one head of A anchors at stride 32;
    ``predict`` is the same decode with a single scale ((1, 160, 5, 30) at 320 x 512)."""

    net_type = "cardensenet"

    def predict(self, batch_out, **kw):
        batch_out = batch_out if isinstance(batch_out, (list, tuple)) else [batch_out]          # car/YOLO.py:897
        return YOLO.predict(self, batch_out, **kw)


class ProjectRectangle6D:
    """yolo_modules/licence_plate_render/__init__.py:274-402 on the GPU: plate corners under a 6-D pose and the perspective crop.
    ``camera`` = dict(image_width, image_height, fx, fy, cx, cy) (the reference reads them from its camera yaml, :279-286)."""

    def __init__(self, camera, device=0):
        self.camera_w, self.camera_h = camera["image_width"], camera["image_height"]
        self.fx, self.fy, self.cx, self.cy = camera["fx"], camera["fy"], camera["cx"], camera["cy"]
        self.device = torch.device("cuda", device)

    def _poses(self, pose_6d):
        p = torch.as_tensor(np.asarray(pose_6d, np.float32) if not torch.is_tensor(pose_6d) else pose_6d, dtype=torch.float32)
        return p.reshape(-1, p.shape[-1]).to(self.device).contiguous()

    def __call__(self, pose_6d):
        """[X, Y, Z (mm), r1, r2, r3 (rad)] -> (4, 2) float32 corner points (:340-353); a (B, 6) batch gives (B, 4, 2)."""
        p = self._poses(pose_6d)
        out = api.lp_corners(p, (self.fx, self.fy, self.cx, self.cy), pose_offset=0).cpu().numpy()
        return out[0] if np.ndim(pose_6d) == 1 else out

    def add_edges(self, img, pose, LP_size=(160, 380)):
        """:379-402: returns (corner points scaled to ``img``, clipped plate).  ``img``: (H, W, 3) uint8 numpy / cuda tensor.  The polyline
        overlay of the reference is drawing, left to the caller (cv2.polylines on the returned corners)."""
        im = torch.as_tensor(img).to(self.device).contiguous()
        p = self._poses(pose)
        corners = api.lp_corners(p[:1], (self.fx, self.fy, self.cx, self.cy), im.shape[1] / float(self.camera_w), im.shape[0] / float(self.camera_h), 0)
        clipped, ok = api.lp_unwarp(im, corners, LP_size)
        return corners[0].cpu().numpy(), clipped[0].cpu().numpy()


def cls2ang(rows, n_class=24):
    """yolo_cv.cls2ang / car/video_node.py:244-252 for predict() rows: -> (vec_ang (B,), vec_rad (B,)) numpy."""
    t = rows if torch.is_tensor(rows) else torch.from_numpy(np.ascontiguousarray(rows, np.float32)).cuda()
    ang, rad = api.azimuth(t.contiguous(), n_class)
    return ang.cpu().numpy(), rad.cpu().numpy()


class CarLPYOLO(YOLO):
    """car_and_LP/YOLO.py ``YOLO``: CarLPNet + predict_LP."""

    net_type = "carlpnet"

    def _train_batch(self, bxs, car_bys, LP_bys=None, car_rotate=False):
        """car_and_LP/YOLO.py:265-304: forward, car targets + losses, LP targets + losses (`_loss_mask_LP`, `_score_weight_LP`,
        `_get_loss_LP`), backward of the sum of all ten, ``trainer.step(batch_size)``.  One list entry per process (torchrun contract).
        ``self.last_losses`` is (10,b): the five car losses then LP_score, LP_xy, LP_z, LP_r, LP_class.  Without ``LP_bys`` it is the
        car-only step of car/YOLO.py:350-399."""
        if LP_bys is None:
            return YOLO._train_batch(self, bxs, car_bys, car_rotate)
        if not hasattr(self, "trainer"):
            self._init_train()
        if len(bxs) != 1 or len(car_bys) != 1 or len(LP_bys) != 1:
            raise ValueError("one process per GPU (torchrun contract, INTEGRATION.md): pass this rank's slice as single-element lists")
        self.last_losses = self.trainer.forward_backward(bxs[0], car_bys[0], self.scale, self.positive_weight, self.negative_weight, car_rotate,
                                                         lp_labels=LP_bys[0], lp_positive_weight=getattr(self, "LP_positive_weight", 1.0),
                                                         lp_negative_weight=getattr(self, "LP_negative_weight", 0.1))
        self.trainer.allreduce_grads()
        self.trainer.step(self.global_batch_size)
        self.backward_counter += 1

    train_step = _train_batch

    def predict_LP(self, LP_batch_out, return_index=False):
        """car_and_LP/YOLO.py:133-157: [LP_x (B,Hs,Ws,10)] -> np.float32 (B,7)."""
        lp = LP_batch_out[0] if isinstance(LP_batch_out, (list, tuple)) else LP_batch_out
        rows, idx = api.decode_lp(lp, 0, self.LP_r_max)
        if return_index:
            return rows.cpu().numpy(), idx.cpu().numpy()
        return rows.cpu().numpy()


class LicencePlateDetectioin:
    """licence_plate/LP_detection.py: DenseNet pose detector."""

    def __init__(self, args=None, spec=None, params=None, precision="fp32", max_batch=1, gpu=None):
        args = _args(args, gpu="0", mode="video")
        spec = _load_spec(args, spec)
        for key in spec:                                   # LP_detection.py:106-107
            setattr(self, key, spec[key])
        self.spec = spec
        self.version = getattr(args, "version", None)
        self.ctx = get_ctx(args.gpu if gpu is None else str(gpu))
        self.num_downsample = len(self.block_config) + 1
        self.net = api.Net("lpdensenet", spec, precision, max_batch, self.ctx[0])
        if params is not None:
            self.net.load_params(params)

    def predict_LP(self, batch_out, return_index=False):
        """LP_detection.py:147-162: NCHW (B,10,H,W) net output -> (10,) for image 0 (argmax of the raw score)."""
        out = batch_out[0] if isinstance(batch_out, (list, tuple)) else batch_out
        rows, idx = api.decode_lp(out, 1, self.LP_r_max)
        rows = rows.cpu().numpy()
        if return_index:
            return rows[0], int(idx[0].item())
        return rows[0]

    def _loss_mask_LP_and_get_loss(self, net_out, LP_by, with_grad=False):
        """GPU replacement of ``_loss_mask_LP`` + ``_get_loss_LP`` (+ the score weights of ``_train_batch_LP``), LP_detection.py:285-360:
        net_out = the NCHW (B,ch,H/32,W/32) network output, LP_by = labels (B,n_obj,>=10).  Returns the (5,B) losses (and the gradient of
        their sum w.r.t. net_out).  The DenseNet's own backward is not part of this build (out of scope for round 2)."""
        out = net_out[0] if isinstance(net_out, (list, tuple)) else net_out
        losses, dlp = api.lp_loss_targets(out, LP_by, 2 ** self.num_downsample, self.LP_r_max, self.scale, self.LP_positive_weight,
                                          self.LP_negative_weight, nchw=True, with_grad=with_grad)
        return (losses, dlp) if with_grad else losses

    def predict_LP_batch(self, batch_out):
        """Same decode applied to every image of the batch -> (B,10) (extension for batched serving)."""
        out = batch_out[0] if isinstance(batch_out, (list, tuple)) else batch_out
        rows, _ = api.decode_lp(out, 1, self.LP_r_max)
        return rows.cpu().numpy()
