// Internal representation of a network handle: parameters, activation buffers, the flat op list.
// Shared by the inference executor (net.cu) and the training step (train_step.cu).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_umma.cuh"

namespace yb {

constexpr float kBnEps = 1e-5f;   // gluon BatchNorm() default (gluoncv _conv2d)

struct Param {
  std::string name;
  int shape[4];
  int ndim;
  std::vector<float> host;
  bool loaded = false;
  size_t numel() const { size_t n = 1; for (int i = 0; i < ndim; ++i) n *= (size_t)shape[i]; return n; }
};

struct Buffer {            // one activation buffer; holds max_batch images
  size_t bytes_per_image = 0;
  size_t offset = 0;       // byte offset inside the workspace
  int dtype = DT_F32;
};

struct View {              // NHWC tensor or a channel slice of a wider buffer (per-image dims)
  int buf = -1;            // >= 0 workspace buffer; -1 network input; <= -2: user output (-2 - index)
  int H = 0, W = 0, C = 0;
  int cpitch = 0, coff = 0;
  int dtype = DT_F32;
  long long ps = 0;        // plane stride in elements (DT_BF16X3: plane p lives at element offset p*ps)
  bool il = false;         // planes interleaved per pixel: row = [plane 0 (C) | plane 1 (C)], cpitch = 2C, ps = C (32-channel fp16x2 tensors)
};

enum OpKind { OP_CONV = 0, OP_POOL = 1, OP_PREACT = 2 };   // PREACT: y = relu(bn(x)) into a zero-padded copy (DenseNet pre-activation)

struct Op {
  int kind = OP_CONV;
  std::string name;        // oracle layer name whose activation this op produces
  View in, out, res;
  bool has_res = false;
  int kh = 1, kw = 1, stride = 1, pad = 0, cout = 0, act = ACT_NONE;
  int upsample2 = 0, out_nchw = 0, is_max = 0;
  int cin_real = 0;        // conv reading a channel-padded pre-activated copy: the convolution's true Cin (weights are zero-padded to in.C)
  int p_weight = -1, p_bias = -1, p_bn = -1, p_prebn = -1;   // index of the FIRST param of each group
  // device parameter pointers (valid after finalize)
  float* w_f32 = nullptr;
  std::vector<float> w_stem_host;   // 3-channel 3x3 stem: host copy of the [27][cout_pad] matrix (kernel-parameter weights)
  float* w_f32_own = nullptr;  // the handle's own copy in the parameter arena (w_f32 points into the trainer's flat buffer while training)
  int cout_pad = 0;
  float *scale = nullptr, *shift = nullptr, *pre_scale = nullptr, *pre_shift = nullptr;
  UmmaConv umma;           // tensor-core path state (tensor maps, packed weights); engaged iff umma.enabled
};

}  // namespace yb

namespace yb { struct TrainState; }
using yb::Param; using yb::Buffer; using yb::View; using yb::Op;

struct yolo_handle {
  yolo_spec spec;
  int device = 0;
  std::string err;
  std::vector<Param> params;
  std::map<std::string, int> pindex;
  std::vector<Buffer> bufs;
  std::vector<Op> ops;
  std::vector<View> outputs;               // per-output view (buf = -2 - i)
  std::map<std::string, View> named;       // oracle layer name -> activation view
  int act_dtype = yb::DT_F32;
  bool finalized = false;
  void* dparams = nullptr;                 // device parameter arena
  size_t dparams_bytes = 0;
  char* ws = nullptr;
  size_t ws_bytes = 0;
  int last_launches = 0;
  double flops_per_image = 0.0;
  // lazily allocated device staging for yolo_predict_host
  void* stage = nullptr;
  size_t stage_bytes = 0;
  int* d_flags = nullptr;                  // device flags: [0] |= 1 activation left the fp16 range of the DT_F16X2 high plane, |= 2 weight did
  yb::TrainState* train = nullptr;         // training state (train_step.cu), null until yolo_train_init
};


namespace yb {
inline int hfail(yolo_handle* h, int code) {
  if (h) h->err = tls_error();
  return code;
}
void train_release(yolo_handle* h, bool writeback = true);      // writeback: copy the trained parameters into the handle's own copies
// debug: the fp32 gradient of a named activation after yolo_train_forward_backward, as NCHW (yolo_debug_activation("grad:<name>"))
int train_debug_grad(yolo_handle* h, const char* layer_name, int batch, float* host_nchw, size_t n_elems);
void fill_conv_desc(const yolo_handle* h, const Op& op, int batch, const void* input, void* const* outputs, ConvDesc& d);
}  // namespace yb
