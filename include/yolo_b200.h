/*
 * yolo_b200.h - C ABI of the B200-native YOLOv3 detection hot path (drop-in for the MXNet/Gluon
 * path of n8886919/YOLO).  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Every entry point returns an int status (0 = OK, negative = YOLO_E_*), never aborts; the text of
 * the last failure on a handle is available through yolo_last_error().  All device work is
 * asynchronous on the cudaStream_t passed as `stream` (void* so the header needs no CUDA include).
 * A handle is not internally locked: "forward on thread A, decode of a copied output on thread B"
 * is safe (reference threading contract, car/video_node.py:131-177); concurrent calls on the SAME
 * handle must be serialised by the caller.
 *
 * Reference interfaces replaced (paths under the reference repo):
 *   yolo_create / yolo_destroy      <- YOLO.__init__ + _init_net      car/YOLO.py:49-110,
 *                                      CarLPNet                       car_and_LP/YOLO.py:47-61,
 *                                      LPDenseNet                     licence_plate/LP_detection.py:59-97
 *   yolo_param_* / yolo_load_param  <- yolo_gluon.init_NN             yolo_modules/yolo_gluon.py:172-201
 *                                      yolo_gluon.init_executor       yolo_modules/yolo_gluon.py:204-242
 *   yolo_forward                    <- net.forward(is_train=False, data=nd_img)
 *                                      car/video_node.py:230-231, licence_plate/LPD_video_node.py:78-79
 *   yolo_decode_top1                <- YOLO.predict                   car/YOLO.py:568-597
 *                                      (+ _init_syxhw :123-155, _yxhw_to_ltrb :552-566, merge_and_slice :841-849)
 *   yolo_decode_nms                 <- north-star extension (the reference has no NMS);
 *                                      IoU form of yolo_gluon.get_iou yolo_modules/yolo_gluon.py:158-167
 *   yolo_decode_lp                  <- YOLO.predict_LP                car_and_LP/YOLO.py:133-169 (mode 0)
 *                                      LicencePlateDetectioin.predict_LP licence_plate/LP_detection.py:147-162 (mode 1)
 *   yolo_loss_targets               <- _loss_mask + _find_best + _score_weight + _get_loss (+ the head gradient of
 *                                      sum(losses).backward())       car/YOLO.py:385-394,401-498, yolo_gluon.get_iou :127-168
 *   yolo_train_init / yolo_train_forward_backward / yolo_train_apply
 *                                   <- _init_train + _train_batch + gluon.Trainer.step(batch_size) with Adam
 *                                      car/YOLO.py:157-207,350-399
 *   yolo_nccl_unique_id / yolo_train_comm_init
 *                                   <- kvstore 'device' gradient reduction inside trainer.step  car/YOLO.py:396
 *   yolo_get_param                  <- net.collect_params().save(...)  car/YOLO.py:546-549 (read-back for checkpoints)
 *   yolo_resize_u8                  <- cv2.resize(img, size) in front of the network   car/video_node.py:150
 *   yolo_azimuth                    <- softmax + atan2 of the orientation classes   car/video_node.py:244-252, yolo_cv.py:85-94
 *   yolo_lp_corners / yolo_lp_unwarp <- ProjectRectangle6D.__call__ / add_edges    yolo_modules/licence_plate_render/__init__.py:340-402
 *   yolo_lp_loss_targets            <- _find_best_LP + _loss_mask_LP + _get_loss_LP licence_plate/LP_detection.py:259-313,354-360
 *   yolo_predict_host               <- cv_img_2_ndarray + net.forward + predict + asnumpy
 *                                      yolo_modules/yolo_gluon.py:335-357, car/YOLO.py:597
 */
#ifndef YOLO_B200_H_
#define YOLO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YOLO_OK          0
#define YOLO_E_BADARG   -1
#define YOLO_E_SHAPE    -2
#define YOLO_E_CUDA     -3
#define YOLO_E_NCCL     -4
#define YOLO_E_OOM      -5
#define YOLO_E_STATE    -6
#define YOLO_E_UNSUPPORTED -7

#define YOLO_MAX_STAGES  8
#define YOLO_MAX_SCALES  3
#define YOLO_MAX_ANCHORS 8
#define YOLO_MAX_BLOCKS  8

/* net_type */
#define YOLO_NET_CARNET      0   /* BasicYOLONet topology + CarNet forward */
#define YOLO_NET_CARLPNET    1   /* CarNet + licence-plate pose branch */
#define YOLO_NET_LPDENSENET  2   /* DenseNet licence-plate detector */
#define YOLO_NET_CARDENSENET 4   /* CarDenseNet (car/utils.py:48-61): the DenseNet with one YOLO head, NHWC output (B, H*W, A, C) */
#define YOLO_NET_DEBUGCONV   3   /* kernel unit-test harness: conv "pre" (3 -> channels[0], 3x3) then conv "test"
                                    (channels[0] -> channels[1], k=layers[0], stride=layers[1], pad=layers[2],
                                    act=layers[3], residual=layers[4], bn=layers[5]); output fp32 NHWC */

/* precision: arithmetic of the convolution path */
#define YOLO_PREC_FP32    0   /* fp32 FFMA implicit GEMM (parity grade)            */
#define YOLO_PREC_BF16    1   /* bf16 operands, fp32 accumulate, tcgen05 tensor cores */
#define YOLO_PREC_BF16X6  2   /* fp32 emulated by 6 bf16 tcgen05 passes on 3-way split operands (24-bit operands) */
#define YOLO_PREC_FP16X3  3   /* fp32 emulated by 3 fp16 tcgen05 passes on 2-way split operands (22-bit operands) */

/* input layouts accepted by yolo_forward */
#define YOLO_IN_NCHW_F32  0   /* (B,3,H,W) fp32 in [0,1] - what cv_img_2_ndarray produces */
#define YOLO_IN_NHWC_U8   1   /* (B,H,W,3) uint8 camera frames; /255 fused into the stem  */

typedef struct yolo_handle yolo_handle;

/* Mirrors the reference's spec.yaml schema (car/v1/spec.yaml, licence_plate/v2/spec.yaml). */
typedef struct yolo_spec {
  int32_t net_type;
  int32_t height, width;                 /* spec `size` */
  int32_t n_layers;                      /* len(spec `layers`)  */
  int32_t layers[YOLO_MAX_STAGES];
  int32_t channels[YOLO_MAX_STAGES + 1]; /* len = n_layers + 1  */
  int32_t n_scales, n_anchors;           /* spec `all_anchors` is (n_scales, n_anchors, 2) as (h, w) */
  float   anchors[YOLO_MAX_SCALES][YOLO_MAX_ANCHORS][2];
  int32_t channels_per_anchor;           /* slice_point[-1]; slices are fixed at [1,3,5,6,C] */
  int32_t lp_channels;                   /* LP_slice_point[-1] (10) */
  float   lp_r_max[3];
  int32_t lp_num_class;
  /* LPDenseNet */
  int32_t num_init_features, growth_rate, bn_size;
  int32_t n_blocks;
  int32_t block_config[YOLO_MAX_BLOCKS];
  /* execution */
  int32_t precision;
  int32_t max_batch;
} yolo_spec;

/* Geometry needed by the decode kernels; independent of any network handle. */
typedef struct yolo_decode_geom {
  int32_t height, width;
  int32_t n_scales, n_anchors, channels_per_anchor;
  int32_t step[YOLO_MAX_SCALES];         /* stride of each scale, shallow -> deep */
  float   anchors[YOLO_MAX_SCALES][YOLO_MAX_ANCHORS][2];
} yolo_decode_geom;

typedef struct yolo_nms_params {
  float   score_thr;   /* keep candidates with sigmoid(score) > score_thr */
  float   iou_thr;     /* suppress same-class boxes with IoU > iou_thr   */
  int32_t max_out;     /* rows written per image                          */
  int32_t max_cand;    /* candidates entering suppression (<= 1024)       */
} yolo_nms_params;

/* Loss hyper-parameters: spec.yaml `scale`, `positive_weight`, `negative_weight` (car/v1/spec.yaml:27-33). */
typedef struct yolo_loss_params {
  float   scale_score, scale_box_yx, scale_box_hw, scale_rotate, scale_class;
  float   positive_weight, negative_weight;
  int32_t car_rotate;   /* 0: rotate loss weight forced to 0 (car/YOLO.py:492) */
} yolo_loss_params;

const char* yolo_version(void);

int  yolo_create(const yolo_spec* spec, int device, yolo_handle** out);
int  yolo_destroy(yolo_handle* h);
const char* yolo_last_error(const yolo_handle* h);   /* h may be NULL: last create() failure */

/* Parameters.  Enumerate in canonical order; shapes are (O,I,kh,kw) for conv weights, (C) otherwise. */
int  yolo_param_count(const yolo_handle* h);
int  yolo_param_info(const yolo_handle* h, int index, const char** name, int32_t shape[4], int32_t* ndim);
int  yolo_load_param(yolo_handle* h, const char* name, const float* host, size_t n_elems);
int  yolo_finalize_params(yolo_handle* h, void* stream);   /* fold BN, repack, upload; requires all params */

/* Activation workspace: caller-owned device memory (e.g. a torch uint8 tensor).  Besides it the handle owns, on its device, the
 * folded parameters and packed weight planes and - allocated on the first forward that needs them, released by yolo_destroy - small
 * per-layer scratch buffers for the split-K partial tiles of the convolution's last wave (a few MB per layer). */
size_t yolo_workspace_bytes(const yolo_handle* h, int batch);
int  yolo_set_workspace(yolo_handle* h, void* device_ptr, size_t bytes);

/* Outputs of forward: n_out tensors; shape query returns per-image dims. */
int  yolo_output_count(const yolo_handle* h);
int  yolo_output_shape(const yolo_handle* h, int index, int32_t shape[4], int32_t* ndim); /* without batch dim */

/* Forward.  `input` and every outputs[i] are device pointers; outputs are fp32:
 *   CARNET     : n_scales heads (B, H_s*W_s, A, C) shallow -> deep
 *   CARLPNET   : the same + LP map (B, H_0, W_0, lp_channels)
 *   LPDENSENET : NCHW (B, 7+lp_num_class, H/32, W/32)                                            */
int  yolo_forward(yolo_handle* h, const void* input, int batch, int in_layout,
                  void* const* outputs, void* stream);

/* fp16x3 range check.  The 16-bit activation format of YOLO_PREC_FP16X3 stores v = hi + lo with an fp16 high plane: |v| > 65504
 * saturates.  Saturation is never silent: every kernel that writes the format ORs a bit into a device flag word of the handle
 * (which bit tells where).  Reads AND CLEARS the flags; synchronises `stream`. */
#define YOLO_SAT_ACT_CONV  1   /* activation written by a tensor-core convolution epilogue            */
#define YOLO_SAT_WEIGHT    2   /* weight re-packed by the training step                                */
#define YOLO_SAT_ACT_FFMA  4   /* activation written by an FFMA convolution / stem / pool              */
#define YOLO_SAT_ACT_BN    8   /* activation written by the training step's normalise pass             */
#define YOLO_SAT_GRAD     16   /* scaled pre-activation gradient written by the BatchNorm backward     */
int  yolo_check_saturation(yolo_handle* h, int32_t* flags_out, void* stream);

/* Debug/parity: copy an internal activation (by oracle layer name) to host as NCHW fp32. */
int  yolo_debug_activation(yolo_handle* h, const char* layer_name, int batch, float* host_nchw, size_t n_elems);

/* Debug/parity: weight gradient of ONE convolution through the tcgen05 weight-gradient kernel of the training step.
 * x: device fp32 NHWC (n,h,w,cin); dz: device fp32 (n*ho*wo, cout); dW: device fp32 [k*k*cin][cout], row = (r*k+s)*cin + c.
 * cin % 64 == 0 and cout % 64 == 0.  variant = 0 (other values probe descriptor conventions in the unit test). */
int  yolo_debug_wgrad(const float* x, const float* dz, int n, int h, int w, int cin, int cout, int k, int stride, int pad,
                      float* dW, int variant, void* stream);

/* Decode + selection (fused, one launch).  heads: device fp32, shallow -> deep.
 * out_rows (B, C) fp32 = [sigmoid(score), y, x, h, w, rotate, class logits...]; out_idx (B) int32 flat index. */
int  yolo_decode_top1(const yolo_decode_geom* g, const void* const* heads, int batch,
                      float* out_rows, int32_t* out_idx, void* stream);
/* out_rows (B, max_out, C), out_idx (B, max_out), out_count (B). out_count[b] < 0: candidate overflow (>4096). */
int  yolo_decode_nms(const yolo_decode_geom* g, const void* const* heads, int batch,
                     const yolo_nms_params* p, float* out_rows, int32_t* out_idx, int32_t* out_count,
                     void* stream);
/* Licence-plate pose decode.  mode 0: lp (B,Hs,Ws,ch) NHWC, argmax of sigmoid(score), out (B,7);
 * mode 1: lp (B,ch,Hs,Ws) NCHW, argmax of the raw score, out (B,ch).  out_idx (B) may be NULL. */
int  yolo_decode_lp(const void* lp, int batch, int hs, int ws, int ch, int mode, const float r_max[3],
                    float* out_rows, int32_t* out_idx, void* stream);

/* Training targets + losses (+ head gradients), GPU-resident (no per-label host sync).
 * labels: device fp32 (B, n_obj, 6+num_class) rows [cls, y, x, h, w, rotate, class distribution...], cls < 0 = no object.
 * out_losses: device fp32 (5, B) = score, box_yx, box_hw, rotate, class (order of spec `loss_name`).
 * dheads: NULL, or n_scales device fp32 tensors shaped like the heads receiving d(sum of the 5 losses)/d(head).
 * out_assign: NULL or device int32 (B, n_obj): matched flat box index per label (-1 = no object).
 * scratch: device memory of yolo_loss_scratch_bytes(batch, n_obj) bytes. */
size_t yolo_loss_scratch_bytes(int batch, int n_obj);
int  yolo_loss_targets(const yolo_decode_geom* g, const void* const* heads, const float* labels, int batch, int n_obj,
                       const yolo_loss_params* p, void* scratch, float* out_losses, void* const* dheads, int32_t* out_assign,
                       void* stream);

/* Post-decode consumers (device in, device out).
 *   azimuth    : rows (B,row_len) as written by yolo_decode_top1; the last n_class entries are the orientation-class logits.
 *                out_angle[b] = atan2(sum sin_k p_k, sum cos_k p_k), p = softmax, k*360/n_class degrees (car/video_node.py:244-252);
 *                out_radius (may be NULL) = rows[b][0] * |mean vector| (yolo_cv.py:85-94).
 *   lp_corners : poses rows hold [.., X, Y, Z (mm), r1, r2, r3 (rad), ..] starting at pose_offset (1 for yolo_decode_lp rows);
 *                intrinsics = {fx, fy, cx, cy}; out_corners (B,4,2) = plate corners in pixels * (x_scale, y_scale)
 *                (ProjectRectangle6D, yolo_modules/licence_plate_render/__init__.py:340-377).
 *   lp_unwarp  : `add_edges` :379-402 - perspective crop of the frame (uint8 HWC, one frame for all plates or one per plate) to an
 *                out_h x out_w x 3 plate image per set of corners; ok[b] = 0 where the quadrilateral is degenerate (crop zero-filled). */
/* Input stage: cv2.resize(frame, (dst_w, dst_h)) (INTER_LINEAR, OpenCV's fixed-point arithmetic: bit-exact when shrinking) for uint8 HWC
 * device frames (B,src_h,src_w,3) -> (B,dst_h,dst_w,3); the result feeds yolo_forward(YOLO_IN_NHWC_U8) (car/video_node.py:150). */
int  yolo_resize_u8(const unsigned char* src, int batch, int src_h, int src_w, unsigned char* dst, int dst_h, int dst_w, void* stream);
int  yolo_azimuth(const float* rows, int batch, int row_len, int n_class, float* out_angle, float* out_radius, void* stream);
int  yolo_lp_corners(const float* poses, int batch, int pose_stride, int pose_offset, const double intrinsics[4], float x_scale,
                     float y_scale, float* out_corners, void* stream);
int  yolo_lp_unwarp(const unsigned char* img, int batch, int img_is_batched, int h, int w, const float* corners, int out_h, int out_w,
                    unsigned char* out, int32_t* ok, void* stream);

/* Licence-plate pose head: targets + the five LP losses (+ gradient of their sum w.r.t. the map) - `_find_best_LP`, `_loss_mask_LP`,
 * `_get_loss_LP` (licence_plate/LP_detection.py:259-313,354-360; `_score_weight_LP` car_and_LP/YOLO.py:124-131).
 * lp_map: device fp32 (B,hs,ws,ch) NHWC (nchw = 0, CarLPNet) or (B,ch,hs,ws) (nchw = 1, LPDenseNet); channels [score, x, y, z, r1, r2, r3, class...].
 * labels: device fp32 (B, n_obj, n_lab >= 10) rows [flag (<0: none), X, Y, Z (mm), r1, r2, r3 (rad), pixel x, pixel y, ..., class index (last)].
 * step: pixels per cell (2^num_downsample).  out_losses: device (5,B) = LP_score, LP_xy, LP_z, LP_r, LP_class.  dlp: NULL or like lp_map. */
typedef struct yolo_lp_loss_params {
  float scale_score, scale_xy, scale_z, scale_r, scale_class;   /* spec `scale` LP_* entries */
  float positive_weight, negative_weight;                       /* LP_positive_weight, LP_negative_weight */
} yolo_lp_loss_params;
int  yolo_lp_loss_targets(const float* lp_map, int nchw, int batch, int hs, int ws, int ch, int step, const float r_max[3],
                          const float* labels, int n_obj, int n_lab, const yolo_lp_loss_params* p, float* out_losses, float* dlp,
                          void* stream);

/* Training step (CARNET / CARLPNET, YOLO_PREC_FP16X3 = fp32-grade arithmetic on the tensor cores), one process per GPU.  The four
 * flat buffers (parameters, gradients, Adam m, Adam v) are caller-owned device memory of yolo_train_flat_size() floats each.
 * BatchNorm statistics stay local to the GPU (car/YOLO.py:94-96).  Every reduction is fixed-order: a step is bit-reproducible.
 *   forward_backward: train-mode forward, targets + the five losses -> out_losses (device, (5,B)), backward of their sum.  With a
 *                     communicator attached (yolo_train_comm_init) the gradient is summed over ranks bucket by bucket WHILE the
 *                     backward runs (ncclAllReduce on a side stream as each bucket's weight gradients complete); the call's stream
 *                     waits for the last bucket.  The reference sums over contexts in trainer.step via kvstore 'device' (car/YOLO.py:396).
 *                     NCCL failures return YOLO_E_NCCL.  Without a communicator the caller may all-reduce `grads_flat` itself.
 *   apply: w -= lr_t * m / (sqrt(v) + eps) with g = grads * rescale_grad (= 1/global batch), lr_t = lr*sqrt(1-b2^t)/(1-b1^t); then the
 *          fp16 weight planes are re-packed on the device and the inference epilogues refolded (yolo_forward serves the trained net).
 *   set_bn_momentum: running = momentum*running + (1-momentum)*batch statistic; default 0.9 (gluon).  0 makes one forward a
 *          calibration pass (synthetic weights).
 *   nccl_unique_id / train_comm_init: rank 0 creates a 128-byte NCCL id, every rank receives it (any transport) and joins.
 *          bucket_bytes = all-reduce granularity (0: 64 MB). */
size_t yolo_train_flat_size(const yolo_handle* h);
int  yolo_train_init(yolo_handle* h, float* params_flat, float* grads_flat, float* adam_m, float* adam_v, size_t n_flat, void* stream);
int  yolo_train_set_bn_momentum(yolo_handle* h, float momentum);
int  yolo_nccl_unique_id(void* id128);
int  yolo_train_comm_init(yolo_handle* h, const void* id128, int rank, int world, size_t bucket_bytes);
int  yolo_train_forward_backward(yolo_handle* h, const void* input, int in_layout, const float* labels, int batch, int n_obj,
                                 const yolo_loss_params* lp, float* out_losses, void* stream);
/* car_and_LP `_train_batch(bxs, car_bys, LP_bys)` (car_and_LP/YOLO.py:265-304), CARLPNET: the backward starts from the sum of the five car
 * losses AND the five LP losses.  lp_labels (B, n_lp_obj, n_lp_lab) as in yolo_lp_loss_targets; out_losses: device (10,B), car rows first. */
int  yolo_train_forward_backward_lp(yolo_handle* h, const void* input, int in_layout, const float* labels, int batch, int n_obj,
                                    const yolo_loss_params* lp, const float* lp_labels, int n_lp_obj, int n_lp_lab,
                                    const yolo_lp_loss_params* lpp, float* out_losses, void* stream);
int  yolo_train_apply(yolo_handle* h, float lr, float beta1, float beta2, float eps, float rescale_grad, void* stream);
/* Read a parameter / running statistic (want_grad = 0) or its gradient (want_grad = 1) back in yolo_load_param's layout. */
int  yolo_get_param(yolo_handle* h, const char* name, float* host, size_t n_elems, int want_grad);

/* End-to-end convenience with HOST buffers (pinned recommended): H2D, forward, decode_top1, D2H, sync.
 * CARNET / CARLPNET only.  host_rows (B, C) ; host_idx (B) may be NULL. */
int  yolo_predict_host(yolo_handle* h, const void* host_input, int batch, int in_layout,
                       float* host_rows, int32_t* host_idx, void* stream);

/* Introspection used by bench.py: launches issued by the last forward()/decode call, conv FLOPs per image. */
int    yolo_last_launch_count(const yolo_handle* h);
double yolo_conv_flops_per_image(const yolo_handle* h);

#ifdef __cplusplus
}
#endif
#endif  /* YOLO_B200_H_ */
