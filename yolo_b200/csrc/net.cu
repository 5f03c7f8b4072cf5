// Layer plan, parameter store, executor and the C ABI of yolo_b200 (include/yolo_b200.h).
//
// The plan is derived from the reference's spec schema exactly like the reference builds its gluon
// blocks (file:line under the reference repo):
//   BasicYOLONet.__init__   yolo_modules/basic_yolo.py:8-39      stem + stages of residual blocks
//   YOLOPyrmaid             yolo_modules/basic_yolo.py:108-123   detection blocks (full pyramid channel), transitions, outputs
//   CarNet.hybrid_forward   car/utils.py:68-95                   deep -> shallow walk, heads returned shallow -> deep
//   CarLPNet                car_and_LP/YOLO.py:47-95             + 5 chained detection blocks + 1x1 conv on the shallowest concat map
//   LPDenseNet              licence_plate/LP_detection.py:59-97  DenseNet-style pose detector
// but as a flat list of fused NHWC convolution launches: BN/activation/bias/residual live in conv
// epilogues, DenseNet pre-activation BN->ReLU in conv prologues, upsample+concat and DenseNet concat are
// channel-offset writes into a shared buffer, and the YOLOOutput transpose disappears because the NHWC
// result of the 1x1 head conv already is (B, H*W, A, C).
#include <cuda_fp16.h>
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_umma.cuh"
#include "net_internal.cuh"

namespace yb {

thread_local int g_launches = 0;

std::string& tls_error() {
  static thread_local std::string e;
  return e;
}

int ensure_dyn_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  YB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({func, dev})) return YOLO_OK;
  YB_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert({func, dev});
  return YOLO_OK;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  tls_error() = buf;
  return code;
}

}  // namespace yb

using namespace yb;

namespace yb {

// ------------------------------------------------------------------------------------------------
// plan builder
// ------------------------------------------------------------------------------------------------
struct Builder {
  yolo_handle* h;
  explicit Builder(yolo_handle* hh) : h(hh) {}

  int add_param(const std::string& name, std::initializer_list<int> shape) {
    Param p;
    p.name = name;
    p.ndim = (int)shape.size();
    int i = 0;
    for (int s : shape) p.shape[i++] = s;
    for (; i < 4; ++i) p.shape[i] = 1;
    h->pindex[name] = (int)h->params.size();
    h->params.push_back(std::move(p));
    return (int)h->params.size() - 1;
  }
  int add_bn(const std::string& name, int c) {
    int i = add_param(name + ".gamma", {c});
    add_param(name + ".beta", {c});
    add_param(name + ".running_mean", {c});
    add_param(name + ".running_var", {c});
    return i;
  }
  // interleave: both fp16 planes of a pixel in one 2C-element row (plane stride C) - the layout the Cin = 32 tensor-core
  // kernel reads as single 128-byte rows (conv_umma.cu KIND 7); every other kernel just sees (cpitch, plane stride)
  View new_buffer(int H, int W, int C, int dtype, bool interleave = false) {
    Buffer b;
    b.dtype = dtype;
    b.bytes_per_image = (size_t)H * W * C * dtype_bytes_per_elem(dtype);
    h->bufs.push_back(b);
    View v;
    v.buf = (int)h->bufs.size() - 1;
    v.H = H; v.W = W; v.C = C; v.cpitch = C; v.coff = 0; v.dtype = dtype;
    v.ps = (long long)h->spec.max_batch * H * W * C;
    if (interleave) { v.cpitch = 2 * C; v.ps = C; v.il = true; }
    return v;
  }
  static View slice(const View& v, int coff, int c) {
    View s = v;
    s.coff = v.coff + coff;
    s.C = c;
    return s;
  }

  // Generic fused conv.  `dst` (optional) is a pre-made view to write into (channel slice of a concat buffer
  // or a user output); its H/W must match the (possibly upsampled) result.
  View conv(const std::string& act_name, const std::string& wname, const View& in, int cout, int k, int pad, int stride,
            int act, const std::string& bn, const std::string& bias, const std::string& prebn, const View* dst = nullptr,
            const View* res = nullptr, int upsample2 = 0, int out_nchw = 0, int cin_real = 0) {
    Op op;
    op.kind = OP_CONV;
    op.name = act_name;
    op.in = in;
    op.cin_real = cin_real > 0 ? cin_real : in.C;
    op.kh = op.kw = k; op.pad = pad; op.stride = stride; op.cout = cout; op.act = act;
    op.upsample2 = upsample2; op.out_nchw = out_nchw;
    if (!prebn.empty()) op.p_prebn = add_bn(prebn, in.C);
    op.p_weight = add_param(wname + ".weight", {cout, op.cin_real, k, k});
    if (!bias.empty()) op.p_bias = add_param(bias + ".bias", {cout});
    if (!bn.empty()) op.p_bn = add_bn(bn, cout);
    int Ho = (in.H + 2 * pad - k) / stride + 1, Wo = (in.W + 2 * pad - k) / stride + 1;
    int Hd = upsample2 ? 2 * Ho : Ho, Wd = upsample2 ? 2 * Wo : Wo;
    if (dst) {
      op.out = *dst;
      op.out.H = Hd; op.out.W = Wd; op.out.C = cout;
    } else {
      static const char* il_env = getenv("YOLO_B200_C32I");
      const bool il = h->act_dtype == DT_F16X2 && cout == 32 && !upsample2 && !(il_env && il_env[0] == '0');
      op.out = new_buffer(Hd, Wd, cout, h->act_dtype, il);
    }
    if (res) { op.res = *res; op.has_res = true; }
    h->flops_per_image += 2.0 * Ho * Wo * (double)cout * k * k * op.cin_real;
    h->ops.push_back(op);
    if (!act_name.empty() && !upsample2) h->named[act_name] = h->ops.back().out;
    return h->ops.back().out;
  }
  // gluoncv _conv2d: conv(no bias) -> BN -> LeakyReLU(0.1)
  View conv_bn_leaky(const std::string& name, const View& in, int cout, int k, int pad, int stride, const View* dst = nullptr,
                     const View* res = nullptr, int upsample2 = 0) {
    return conv(name, name, in, cout, k, pad, stride, ACT_LEAKY, name, "", "", dst, res, upsample2);
  }
  View pool(const std::string& name, const View& in, int k, int stride, int pad, int is_max, const View* dst = nullptr) {
    Op op;
    op.kind = OP_POOL;
    op.name = name;
    op.in = in;
    op.kh = op.kw = k; op.stride = stride; op.pad = pad; op.is_max = is_max; op.cout = in.C;
    int Ho = (in.H + 2 * pad - k) / stride + 1, Wo = (in.W + 2 * pad - k) / stride + 1;
    if (dst) { op.out = *dst; op.out.H = Ho; op.out.W = Wo; op.out.C = in.C; }
    else op.out = new_buffer(Ho, Wo, in.C, h->act_dtype);
    h->ops.push_back(op);
    if (!name.empty()) h->named[name] = h->ops.back().out;
    return h->ops.back().out;
  }
  // DenseNet pre-activation convolution: BN(prebn) -> ReLU -> conv.  With a 16-bit activation format the BN -> ReLU is materialised by its
  // own pass into `scratch` (channels zero-padded to a multiple of 64) and the convolution runs on the tensor cores with zero-padded
  // weights; otherwise the FFMA kernel applies the prologue while gathering.
  View preact_conv(const std::string& wname, const View& in, View& scratch, int cout, int k, int pad, int act, const std::string& bn,
                   const std::string& bias, const std::string& prebn, const View* dst = nullptr) {
    const bool tc = h->act_dtype != DT_F32 && scratch.buf >= 0;      // the pass itself copes with unaligned concat slices
    if (!tc) return conv("", wname, in, cout, k, pad, 1, act, bn, bias, prebn, dst);
    const int cpad = (in.C + 63) & ~63;
    Op op;
    op.kind = OP_PREACT;
    op.in = in;
    op.out = scratch;
    op.out.C = cpad;
    op.cout = cpad;
    op.p_prebn = add_bn(prebn, in.C);
    h->ops.push_back(op);
    return conv("", wname, h->ops.back().out, cout, k, pad, 1, act, bn, bias, "", dst, nullptr, 0, 0, in.C);
  }

  // gluoncv YOLODetectionBlockV3(c): returns route, sets tip
  View detection_block(const std::string& name, View x, int c, View* tip) {
    for (int i = 0; i < 2; ++i) {
      x = conv_bn_leaky(name + ".body." + std::to_string(2 * i), x, c, 1, 0, 1);
      x = conv_bn_leaky(name + ".body." + std::to_string(2 * i + 1), x, c * 2, 3, 1, 1);
    }
    x = conv_bn_leaky(name + ".body.4", x, c, 1, 0, 1);
    *tip = conv_bn_leaky(name + ".tip", x, c * 2, 3, 1, 1);
    return x;
  }

  int build_yolo(bool lp_branch) {
    const yolo_spec& s = h->spec;
    const int nl = s.n_layers, npyr = s.n_scales, A = s.n_anchors, C = s.channels_per_anchor;
    if (nl < 1 || nl > YOLO_MAX_STAGES || npyr < 1 || npyr > YOLO_MAX_SCALES || npyr > nl)
      return fail(YOLO_E_BADARG, "spec: n_layers=%d n_scales=%d invalid", nl, npyr);
    if (A < 1 || A > YOLO_MAX_ANCHORS || C < 6) return fail(YOLO_E_BADARG, "spec: n_anchors=%d channels_per_anchor=%d invalid", A, C);
    const int down = 1 << nl;
    if (s.height % down || s.width % down)
      return fail(YOLO_E_SHAPE, "spec: size %dx%d must be divisible by 2^n_layers=%d (route/upsample shapes would not match, car/utils.py:93)",
                  s.height, s.width, down);
    for (int i = 0; i <= nl; ++i)
      if (s.channels[i] < 1 || (i > 0 && s.channels[i] % 2)) return fail(YOLO_E_BADARG, "spec: channels[%d]=%d invalid", i, s.channels[i]);

    View in;
    in.buf = -1; in.H = s.height; in.W = s.width; in.C = 3; in.cpitch = 3; in.coff = 0; in.dtype = DT_F32;
    View x = conv_bn_leaky("stages.0", in, s.channels[0], 3, 1, 1);
    std::vector<View> routes;        // shallow -> deep
    std::vector<View> concat;        // concat buffer per non-deepest route (full view)
    const int nstage = nl + 1;
    for (int st = 1; st <= nl; ++st) {
      const int ch = s.channels[st];
      const bool is_route = st >= nstage - npyr;
      const bool is_concat_route = is_route && st != nl;
      View cat, route_dst;
      if (is_concat_route) {
        cat = new_buffer(x.H / 2, x.W / 2, 2 * ch, h->act_dtype);
        route_dst = slice(cat, ch, ch);          // concat(upsample, route): route is the SECOND half
      }
      const int nblk = s.layers[st - 1];
      const std::string sn = "stages." + std::to_string(st);
      x = conv_bn_leaky(sn + ".0", x, ch, 3, 1, 2, (is_concat_route && nblk == 0) ? &route_dst : nullptr);
      for (int j = 1; j <= nblk; ++j) {
        const std::string bn = sn + "." + std::to_string(j);
        View y = conv_bn_leaky(bn + ".body.0", x, ch / 2, 1, 0, 1);
        x = conv_bn_leaky(bn + ".body.1", y, ch, 3, 1, 1, (is_concat_route && j == nblk) ? &route_dst : nullptr, &x);
        h->named.erase(bn + ".body.1");        // the residual add is fused: this activation is the block output
        h->named[bn] = x;
      }
      if (is_route) { routes.push_back(x); concat.push_back(cat); }
    }
    // pyramid, deep -> shallow
    for (int i = 0; i < npyr; ++i) {
      const bool last = i == npyr - 1;
      const int pc = s.channels[nl - i];
      if (last && lp_branch) {
        const int lpc = s.channels[nl - 2 < 0 ? 0 : nl - 2];     // spec['channels'][-3]
        View t = x, tip;
        for (int b = 0; b < 5; ++b) { detection_block("LP_branch." + std::to_string(b), t, lpc, &tip); t = tip; }
        View lpo;
        lpo.buf = -2 - npyr; lpo.H = t.H; lpo.W = t.W; lpo.C = s.lp_channels; lpo.cpitch = s.lp_channels; lpo.coff = 0; lpo.dtype = DT_F32;
        conv("LP_branch.5", "LP_branch.5", t, s.lp_channels, 1, 0, 1, ACT_NONE, "", "LP_branch.5", "", &lpo);
        if ((int)h->outputs.size() < npyr + 1) h->outputs.resize(npyr + 1);
        h->outputs[npyr] = lpo;
      }
      View tip;
      View route = detection_block("yolo_blocks." + std::to_string(i), x, pc, &tip);
      View ho;
      const int oi = npyr - 1 - i;                                 // heads are returned shallow -> deep
      ho.buf = -2 - oi; ho.H = tip.H; ho.W = tip.W; ho.C = A * C; ho.cpitch = A * C; ho.coff = 0; ho.dtype = DT_F32;
      const std::string on = "yolo_outputs." + std::to_string(i);
      conv(on, on, tip, A * C, 1, 0, 1, ACT_NONE, "", on, "", &ho);
      if ((int)h->outputs.size() < npyr) h->outputs.resize(npyr);
      h->outputs[oi] = ho;
      if (last) break;
      const int nc = s.channels[nl - i - 1];
      const View& cat = concat[npyr - 2 - i];
      View up_dst = slice(cat, 0, nc);
      conv_bn_leaky("transitions." + std::to_string(i), route, nc, 1, 0, 1, &up_dst, nullptr, 1);
      x = cat;
    }
    return YOLO_OK;
  }

  // Unit-test harness for the conv kernels (YOLO_NET_DEBUGCONV): "pre" makes an activation tensor in the
  // precision's storage format, "test" is the convolution under test writing fp32 NHWC to the user output.
  int build_debugconv() {
    const yolo_spec& s = h->spec;
    const int cin = s.channels[0], cout = s.channels[1], k = s.layers[0], stride = s.layers[1], pad = s.layers[2];
    const int act = s.layers[3], residual = s.layers[4], bn = s.layers[5];
    // residual: 0 = fp32 output; 1 = + input, stored in the activation format; 2 = activation format without the add
    if (cin < 1 || cout < 1 || k < 1 || stride < 1 || pad < 0 || residual < 0 || residual > 2 ||
        (residual == 1 && (cin != cout || stride != 1 || 2 * pad != k - 1)))
      return fail(YOLO_E_BADARG, "spec: debug conv parameters invalid");
    View in;
    in.buf = -1; in.H = s.height; in.W = s.width; in.C = 3; in.cpitch = 3; in.coff = 0; in.dtype = DT_F32;
    View x = conv_bn_leaky("pre", in, cin, 3, 1, 1);
    View o;
    o.buf = -2; o.H = (x.H + 2 * pad - k) / stride + 1; o.W = (x.W + 2 * pad - k) / stride + 1; o.C = cout; o.cpitch = cout; o.coff = 0;
    o.dtype = DT_F32;
    if (residual) {
      // residual adds need matching storage formats: route through an activation buffer, then a 1x1 "copy" is not
      // exact - instead keep the residual conv in activation format and expose it by name ("test")
      conv("test", "test", x, cout, k, pad, stride, act, bn ? "test" : "", bn ? "" : "test", "", nullptr, residual == 1 ? &x : nullptr);
      View y = h->named["test"];
      // fp32 view of the result for the caller: identity 1x1 conv would round; the test reads "test" by name instead.
      conv("out", "out", y, 8, 1, 0, 1, ACT_NONE, "", "out", "", &o);
      o.C = 8; o.cpitch = 8;
      h->ops.back().out = o;
    } else {
      conv("test", "test", x, cout, k, pad, stride, act, bn ? "test" : "", bn ? "" : "test", "", &o);
    }
    h->outputs.resize(1);
    h->outputs[0] = o;
    return YOLO_OK;
  }

  // dense_yolo: CarDenseNet (car/utils.py:48-61) = the same DenseNet with (7 + classes) = A*C output channels, returned NHWC as
  // (B, H*W, A, C) like a YOLO head (x.transpose((0, 2, 3, 1)).reshape((0, -1, A, C)))
  int build_lpdense(bool dense_yolo = false) {
    const yolo_spec& s = h->spec;
    if (s.n_blocks < 1 || s.n_blocks > YOLO_MAX_BLOCKS || s.num_init_features < 1 || s.growth_rate < 1 || s.bn_size < 1)
      return fail(YOLO_E_BADARG, "spec: DenseNet parameters invalid");
    const int down = 4 << (s.n_blocks - 1);
    if (s.height % down || s.width % down) return fail(YOLO_E_SHAPE, "spec: size %dx%d must be divisible by %d", s.height, s.width, down);
    const int g = s.growth_rate, dt = h->act_dtype;
    View in;
    in.buf = -1; in.H = s.height; in.W = s.width; in.C = 3; in.cpitch = 3; in.coff = 0; in.dtype = DT_F32;
    View x = conv("stem.bn", "stem.conv", in, s.num_init_features, 7, 3, 2, ACT_RELU, "stem.bn", "", "");
    int nf = s.num_init_features;
    View blk = new_buffer((x.H + 2 - 3) / 2 + 1, (x.W + 2 - 3) / 2 + 1, nf + s.block_config[0] * g, dt);
    View dst = slice(blk, 0, nf);
    pool("stem.pool", x, 3, 2, 1, 1, &dst);
    auto scratch_for = [&](const View& like, int cmax) {             // pre-activated copy of the widest layer input at this resolution
      View none;
      if (dt == DT_F32) return none;
      return new_buffer(like.H, like.W, (cmax + 63) & ~63, dt);
    };
    for (int b = 1; b <= s.n_blocks; ++b) {
      const int nlay = s.block_config[b - 1];
      int cur = nf;
      View scratch = scratch_for(blk, nf + nlay * g);
      for (int l = 0; l < nlay; ++l) {
        const std::string p = "block" + std::to_string(b) + ".layer" + std::to_string(l);
        View xin = slice(blk, 0, cur);
        View y = preact_conv(p + ".conv1", xin, scratch, s.bn_size * g, 1, 0, ACT_RELU, p + ".bn2", "", p + ".bn1");
        View d2 = slice(blk, cur, g);
        conv("", p + ".conv2", y, g, 3, 1, 1, ACT_NONE, "", "", "", &d2);
        cur += g;
        h->named[p] = slice(blk, 0, cur);
      }
      nf = cur;
      if (b != s.n_blocks) {
        const std::string p = "trans" + std::to_string(b);
        View xin = slice(blk, 0, nf);
        View y = preact_conv(p + ".conv", xin, scratch, nf / 2, 1, 0, ACT_NONE, "", "", p + ".bn");
        nf = nf / 2;
        View nb = new_buffer(y.H / 2, y.W / 2, nf + s.block_config[b] * g, dt);
        View d2 = slice(nb, 0, nf);
        pool(p, y, 2, 2, 0, 0, &d2);
        blk = nb;
      } else {
        View xin = slice(blk, 0, nf);
        View t = preact_conv("tail.conv1", xin, scratch, 512, 3, 1, ACT_RELU, "tail.bn2", "tail.conv1", "tail.bn1");
        blk = t;                                                       // carried out of the loop below
      }
    }
    View t = blk;
    const int nout = dense_yolo ? s.n_anchors * s.channels_per_anchor : 7 + s.lp_num_class;
    if (dense_yolo && (s.n_anchors < 1 || s.n_anchors > YOLO_MAX_ANCHORS || s.channels_per_anchor < 6)) return fail(YOLO_E_BADARG, "spec: CarDenseNet anchors/channels invalid");
    View o;
    o.buf = -2; o.H = t.H; o.W = t.W; o.C = nout; o.cpitch = nout; o.coff = 0; o.dtype = DT_F32;
    conv("tail.conv2", "tail.conv2", t, nout, 1, 0, 1, ACT_NONE, "", "tail.conv2", "", &o, nullptr, 0, /*out_nchw=*/dense_yolo ? 0 : 1);
    h->outputs.resize(1);
    h->outputs[0] = o;
    return YOLO_OK;
  }
};

static void layout_workspace(yolo_handle* h) {
  size_t off = 0;
  for (auto& b : h->bufs) {
    b.offset = off;
    size_t bytes = b.bytes_per_image * (size_t)h->spec.max_batch;
    off += (bytes + 1023) & ~(size_t)1023;     // 1 KB alignment keeps every buffer TMA/vector friendly
  }
  h->ws_bytes = off;
}

static void* resolve(const yolo_handle* h, const View& v, const void* input, void* const* outputs) {
  if (v.buf >= 0) return h->ws + h->bufs[v.buf].offset;
  if (v.buf == -1) return const_cast<void*>(input);
  return outputs ? outputs[-2 - v.buf] : nullptr;
}

// ConvDesc of one planned convolution for `batch` images (inference epilogue: folded BN, activation, residual, placement)
void fill_conv_desc(const yolo_handle* h, const Op& op, int batch, const void* input, void* const* outputs, ConvDesc& d) {
  memset(&d, 0, sizeof(d));
  d.in = resolve(h, op.in, input, outputs);
  d.in_dtype = op.in.dtype; d.N = batch; d.H = op.in.H; d.W = op.in.W; d.Cin = op.in.C;
  d.in_cpitch = op.in.cpitch; d.in_coff = op.in.coff; d.in_plane_stride = op.in.ps;
  d.kh = op.kh; d.kw = op.kw; d.stride = op.stride; d.pad = op.pad; d.Cout = op.cout;
  d.w_f32 = op.w_f32; d.cout_pad = op.cout_pad;
  d.w_host = (op.w_f32 == op.w_f32_own && !op.w_stem_host.empty()) ? op.w_stem_host.data() : nullptr;     // stale while a trainer owns the weights
  d.pre_scale = op.pre_scale; d.pre_shift = op.pre_shift;
  d.scale = op.scale; d.shift = op.shift; d.act = op.act;
  if (op.has_res) { d.res = resolve(h, op.res, input, outputs); d.res_cpitch = op.res.cpitch; d.res_coff = op.res.coff; d.res_plane_stride = op.res.ps; }
  d.out = resolve(h, op.out, input, outputs);
  d.out_dtype = op.out.dtype;
  d.Ho = (op.in.H + 2 * op.pad - op.kh) / op.stride + 1;
  d.Wo = (op.in.W + 2 * op.pad - op.kw) / op.stride + 1;
  d.out_cpitch = op.out.cpitch; d.out_coff = op.out.coff; d.out_plane_stride = op.out.ps;
  d.upsample2 = op.upsample2; d.out_nchw = op.out_nchw;
  d.sat_flag = h->d_flags;
}

}  // namespace yb

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" const char* yolo_version(void) { return "yolo_b200 0.1 (sm_100a)"; }

extern "C" const char* yolo_last_error(const yolo_handle* h) { return h ? h->err.c_str() : tls_error().c_str(); }

extern "C" int yolo_create(const yolo_spec* spec, int device, yolo_handle** out) {
  if (!spec || !out) return fail(YOLO_E_BADARG, "create: null spec/out");
  *out = nullptr;
  if (spec->max_batch < 1) return fail(YOLO_E_BADARG, "create: max_batch=%d", spec->max_batch);
  if (spec->precision < YOLO_PREC_FP32 || spec->precision > YOLO_PREC_FP16X3) return fail(YOLO_E_BADARG, "create: precision=%d", spec->precision);
  std::unique_ptr<yolo_handle> h(new yolo_handle());
  h->spec = *spec;
  h->device = device;
  h->act_dtype = spec->precision == YOLO_PREC_BF16 ? DT_BF16 : (spec->precision == YOLO_PREC_BF16X6 ? DT_BF16X3 : (spec->precision == YOLO_PREC_FP16X3 ? DT_F16X2 : DT_F32));
  Builder b(h.get());
  int rc;
  switch (spec->net_type) {
    case YOLO_NET_CARNET: rc = b.build_yolo(false); break;
    case YOLO_NET_CARLPNET:
      if (spec->lp_channels < 7 || spec->n_layers < 3) return fail(YOLO_E_BADARG, "create: CARLPNET needs lp_channels >= 7 and >= 3 stages");
      rc = b.build_yolo(true);
      break;
    case YOLO_NET_LPDENSENET: rc = b.build_lpdense(); break;
    case YOLO_NET_CARDENSENET: rc = b.build_lpdense(true); break;
    case YOLO_NET_DEBUGCONV: rc = b.build_debugconv(); break;
    default: return fail(YOLO_E_BADARG, "create: net_type=%d", spec->net_type);
  }
  if (rc) return rc;
  layout_workspace(h.get());
  *out = h.release();
  return YOLO_OK;
}

extern "C" int yolo_destroy(yolo_handle* h) {
  if (!h) return YOLO_OK;
  cudaSetDevice(h->device);
  if (h->dparams) cudaFree(h->dparams);
  if (h->stage) cudaFree(h->stage);
  if (h->d_flags) cudaFree(h->d_flags);
  train_release(h, false);
  for (auto& op : h->ops) umma_release(op.umma);
  delete h;
  return YOLO_OK;
}

extern "C" int yolo_param_count(const yolo_handle* h) { return h ? (int)h->params.size() : YOLO_E_BADARG; }

extern "C" int yolo_param_info(const yolo_handle* h, int index, const char** name, int32_t shape[4], int32_t* ndim) {
  if (!h || index < 0 || index >= (int)h->params.size()) return fail(YOLO_E_BADARG, "param_info: bad index");
  const Param& p = h->params[index];
  if (name) *name = p.name.c_str();
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = p.shape[i];
  if (ndim) *ndim = p.ndim;
  return YOLO_OK;
}

extern "C" int yolo_load_param(yolo_handle* h, const char* name, const float* host, size_t n_elems) {
  if (!h || !name || !host) return fail(YOLO_E_BADARG, "load_param: null argument");
  auto it = h->pindex.find(name);
  if (it == h->pindex.end()) return hfail(h, fail(YOLO_E_BADARG, "load_param: unknown parameter '%s'", name));
  Param& p = h->params[it->second];
  if (p.numel() != n_elems) return hfail(h, fail(YOLO_E_SHAPE, "load_param: '%s' expects %zu elements, got %zu", name, p.numel(), n_elems));
  p.host.assign(host, host + n_elems);
  p.loaded = true;
  h->finalized = false;
  return YOLO_OK;
}

namespace yb {
static void bn_fold(const yolo_handle* h, int p_bn, int c, std::vector<float>& scale, std::vector<float>& shift) {
  const float* g = h->params[p_bn].host.data();
  const float* b = h->params[p_bn + 1].host.data();
  const float* m = h->params[p_bn + 2].host.data();
  const float* v = h->params[p_bn + 3].host.data();
  scale.resize(c); shift.resize(c);
  for (int i = 0; i < c; ++i) {
    float s = g[i] / sqrtf(v[i] + kBnEps);
    scale[i] = s;
    shift[i] = b[i] - m[i] * s;
  }
}
}  // namespace yb

extern "C" int yolo_finalize_params(yolo_handle* h, void* stream) {
  if (!h) return fail(YOLO_E_BADARG, "finalize: null handle");
  for (auto& p : h->params)
    if (!p.loaded) return hfail(h, fail(YOLO_E_STATE, "finalize: parameter '%s' was never loaded", p.name.c_str()));
  YB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  train_release(h, false);                   // a trainer attached to the previous parameters is stale
  // size the device arena
  size_t total = 0;
  auto take = [&](size_t nfloat) { size_t o = total; total += ((nfloat * 4 + 255) & ~(size_t)255); return o; };
  struct Slot { size_t w, sc, sh, ps, pb; };
  std::vector<Slot> slots(h->ops.size());
  for (size_t i = 0; i < h->ops.size(); ++i) {
    Op& op = h->ops[i];
    if (op.kind == OP_PREACT) { slots[i].ps = take(op.in.C); slots[i].pb = take(op.in.C); continue; }
    if (op.kind != OP_CONV) continue;
    op.cout_pad = (op.cout + 3) & ~3;
    size_t K = (size_t)op.kh * op.kw * op.in.C;
    slots[i].w = take(K * op.cout_pad);
    slots[i].sc = take(op.cout);
    slots[i].sh = take(op.cout);
    slots[i].ps = take(op.in.C);
    slots[i].pb = take(op.in.C);
  }
  if (h->dparams && h->dparams_bytes < total) { cudaFree(h->dparams); h->dparams = nullptr; }
  if (!h->dparams) {
    if (cudaMalloc(&h->dparams, total) != cudaSuccess) { cudaGetLastError(); return hfail(h, fail(YOLO_E_OOM, "finalize: cudaMalloc(%zu) failed", total)); }
    h->dparams_bytes = total;
  }
  if (!h->d_flags) {
    YB_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_flags), 64));
    YB_CUDA(cudaMemsetAsync(h->d_flags, 0, 64, st));
  }
  std::vector<float> stagev(total / 4, 0.f);
  char* dbase = static_cast<char*>(h->dparams);
  for (size_t i = 0; i < h->ops.size(); ++i) {
    Op& op = h->ops[i];
    if (op.kind == OP_PREACT) {                // folded BatchNorm of the DenseNet pre-activation pass
      std::vector<float> sc, sh;
      bn_fold(h, op.p_prebn, op.in.C, sc, sh);
      memcpy(stagev.data() + slots[i].ps / 4, sc.data(), op.in.C * 4);
      memcpy(stagev.data() + slots[i].pb / 4, sh.data(), op.in.C * 4);
      op.pre_scale = reinterpret_cast<float*>(dbase + slots[i].ps);
      op.pre_shift = reinterpret_cast<float*>(dbase + slots[i].pb);
      continue;
    }
    if (op.kind != OP_CONV) continue;
    // cin = channels of the tensor the kernel reads; a convolution behind a pre-activation pass reads a channel-padded copy: its
    // weights are zero-padded from cin_real to cin
    const int cin = op.in.C, creal = op.cin_real, cout = op.cout, kh = op.kh, kw = op.kw;
    const float* W = h->params[op.p_weight].host.data();     // OIHW, I = creal
    float* wd = stagev.data() + slots[i].w / 4;
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < creal; ++c)
        for (int r = 0; r < kh; ++r)
          for (int s2 = 0; s2 < kw; ++s2)
            wd[((size_t)(r * kw + s2) * cin + c) * op.cout_pad + o] = W[(((size_t)o * creal + c) * kh + r) * kw + s2];
    op.w_f32 = op.w_f32_own = reinterpret_cast<float*>(dbase + slots[i].w);
    op.w_stem_host.clear();
    if (cin == 3 && ((kh == 3 && kw == 3 && cout <= 32) || (kh == 7 && kw == 7 && cout <= 64)) && op.cout_pad == cout)
      op.w_stem_host.assign(wd, wd + (size_t)kh * kw * 3 * op.cout_pad);
    std::vector<float> sc, sh;
    op.scale = op.shift = op.pre_scale = op.pre_shift = nullptr;
    if (op.p_bn >= 0) {
      bn_fold(h, op.p_bn, cout, sc, sh);
      if (op.p_bias >= 0) {                    // y = bn(conv + bias)
        const float* bias = h->params[op.p_bias].host.data();
        for (int o = 0; o < cout; ++o) sh[o] += bias[o] * sc[o];
      }
    } else if (op.p_bias >= 0) {
      sc.assign(cout, 1.f);
      sh.assign(h->params[op.p_bias].host.begin(), h->params[op.p_bias].host.end());
    }
    if (!sc.empty()) {
      memcpy(stagev.data() + slots[i].sc / 4, sc.data(), cout * 4);
      memcpy(stagev.data() + slots[i].sh / 4, sh.data(), cout * 4);
      op.scale = reinterpret_cast<float*>(dbase + slots[i].sc);
      op.shift = reinterpret_cast<float*>(dbase + slots[i].sh);
    }
    if (op.p_prebn >= 0) {
      bn_fold(h, op.p_prebn, cin, sc, sh);
      memcpy(stagev.data() + slots[i].ps / 4, sc.data(), cin * 4);
      memcpy(stagev.data() + slots[i].pb / 4, sh.data(), cin * 4);
      op.pre_scale = reinterpret_cast<float*>(dbase + slots[i].ps);
      op.pre_shift = reinterpret_cast<float*>(dbase + slots[i].pb);
    }
  }
  YB_CUDA(cudaMemcpyAsync(h->dparams, stagev.data(), total, cudaMemcpyHostToDevice, st));
  YB_CUDA(cudaStreamSynchronize(st));
  // tensor-core path: pack bf16 weights (needs the host weights) - maps are built in set_workspace
  for (auto& op : h->ops) {
    if (op.kind != OP_CONV) continue;
    const float* W = h->params[op.p_weight].host.data();
    std::vector<float> wpad;
    if (op.cin_real != op.in.C) {              // zero-pad the input-channel dimension
      wpad.assign((size_t)op.cout * op.in.C * op.kh * op.kw, 0.f);
      const size_t khw = (size_t)op.kh * op.kw;
      for (int o = 0; o < op.cout; ++o)
        memcpy(&wpad[(size_t)o * op.in.C * khw], W + (size_t)o * op.cin_real * khw, (size_t)op.cin_real * khw * 4);
      W = wpad.data();
    }
    int rc = umma_prepare_weights(op.umma, h->spec.precision, W, op.cout, op.in.C, op.kh, op.kw,
                                  op.stride, op.pad, op.in.dtype, op.pre_scale != nullptr, op.out_nchw, op.in.il, st);
    if (rc) return hfail(h, rc);
  }
  h->finalized = true;
  if (h->ws) return yolo_set_workspace(h, h->ws, h->ws_bytes);
  return YOLO_OK;
}

extern "C" size_t yolo_workspace_bytes(const yolo_handle* h, int batch) {
  if (!h || batch > h->spec.max_batch) return 0;
  return h->ws_bytes;
}

extern "C" int yolo_set_workspace(yolo_handle* h, void* device_ptr, size_t bytes) {
  if (!h || !device_ptr) return fail(YOLO_E_BADARG, "set_workspace: null argument");
  if (bytes < h->ws_bytes) return hfail(h, fail(YOLO_E_BADARG, "set_workspace: %zu bytes given, %zu required", bytes, h->ws_bytes));
  if (reinterpret_cast<uintptr_t>(device_ptr) & 1023) return hfail(h, fail(YOLO_E_BADARG, "set_workspace: pointer must be 1024-byte aligned"));
  h->ws = static_cast<char*>(device_ptr);
  if (h->finalized) {
    for (auto& op : h->ops) {
      if (op.kind != OP_CONV || !op.umma.eligible) continue;
      if (op.in.buf < 0) { op.umma.enabled = false; continue; }
      // epilogue constraints of the tensor-core kernel: 16-byte aligned 16-bit channel slices, 32-column residual chunks
      const bool out16 = op.out.dtype != DT_F32;
      if ((out16 && (((op.out.cpitch | op.out.coff) & 7) || op.cout % 8)) || (op.has_res && (((op.res.cpitch | op.res.coff) & 7) || op.cout % 32 || !out16))) {
        op.umma.enabled = false;
        continue;
      }
      int rc = umma_build_maps(op.umma, h->ws + h->bufs[op.in.buf].offset, h->spec.max_batch, op.in.H, op.in.W, op.in.C,
                               op.in.cpitch, op.in.coff);
      if (rc) return hfail(h, rc);
    }
  }
  return YOLO_OK;
}

extern "C" int yolo_output_count(const yolo_handle* h) { return h ? (int)h->outputs.size() : YOLO_E_BADARG; }

extern "C" int yolo_output_shape(const yolo_handle* h, int index, int32_t shape[4], int32_t* ndim) {
  if (!h || index < 0 || index >= (int)h->outputs.size() || !shape || !ndim) return fail(YOLO_E_BADARG, "output_shape: bad argument");
  const View& v = h->outputs[index];
  const yolo_spec& s = h->spec;
  if (s.net_type == YOLO_NET_LPDENSENET) {
    shape[0] = v.C; shape[1] = v.H; shape[2] = v.W; shape[3] = 1; *ndim = 3;
  } else if (s.net_type == YOLO_NET_CARDENSENET) {
    shape[0] = v.H * v.W; shape[1] = s.n_anchors; shape[2] = s.channels_per_anchor; shape[3] = 1; *ndim = 3;
  } else if (s.net_type == YOLO_NET_DEBUGCONV) {
    shape[0] = v.H; shape[1] = v.W; shape[2] = v.C; shape[3] = 1; *ndim = 3;
  } else if (index < s.n_scales) {
    shape[0] = v.H * v.W; shape[1] = s.n_anchors; shape[2] = s.channels_per_anchor; shape[3] = 1; *ndim = 3;
  } else {
    shape[0] = v.H; shape[1] = v.W; shape[2] = v.C; shape[3] = 1; *ndim = 3;
  }
  return YOLO_OK;
}

extern "C" int yolo_forward(yolo_handle* h, const void* input, int batch, int in_layout, void* const* outputs, void* stream) {
  if (!h || !input || !outputs) return fail(YOLO_E_BADARG, "forward: null argument");
  if (!h->finalized) return hfail(h, fail(YOLO_E_STATE, "forward: parameters not finalized"));
  if (!h->ws) return hfail(h, fail(YOLO_E_STATE, "forward: workspace not set"));
  if (batch < 1 || batch > h->spec.max_batch) return hfail(h, fail(YOLO_E_SHAPE, "forward: batch=%d outside [1,%d]", batch, h->spec.max_batch));
  if (in_layout != YOLO_IN_NCHW_F32 && in_layout != YOLO_IN_NHWC_U8) return hfail(h, fail(YOLO_E_BADARG, "forward: in_layout=%d", in_layout));
  for (size_t i = 0; i < h->outputs.size(); ++i)
    if (!outputs[i]) return hfail(h, fail(YOLO_E_BADARG, "forward: outputs[%zu] is null", i));
  cudaError_t ce = cudaSetDevice(h->device);
  if (ce != cudaSuccess) return hfail(h, fail(YOLO_E_CUDA, "forward: cudaSetDevice(%d): %s", h->device, cudaGetErrorString(ce)));
  {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, input) == cudaSuccess && pa.type == cudaMemoryTypeDevice && pa.device != h->device)
      return hfail(h, fail(YOLO_E_BADARG, "forward: input lives on device %d but the handle was created for device %d", pa.device, h->device));
    cudaGetLastError();
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int launches0 = g_launches;
  for (auto& op : h->ops) {
    int rc;
    if (op.kind == OP_PREACT) {
      rc = launch_preact(resolve(h, op.in, input, outputs), resolve(h, op.out, input, outputs), op.in.dtype, (long long)batch * op.in.H * op.in.W,
                         op.in.C, op.out.C, op.in.cpitch, op.in.coff, op.in.ps, op.out.cpitch, op.out.ps, op.pre_scale, op.pre_shift, h->d_flags, st);
    } else if (op.kind == OP_POOL) {
      rc = launch_pool(resolve(h, op.in, input, outputs), resolve(h, op.out, input, outputs), op.in.dtype, batch, op.in.H, op.in.W,
                       op.in.C, op.in.cpitch, op.in.coff, op.in.ps, op.out.cpitch, op.out.coff, op.out.ps, op.kh, op.stride, op.pad,
                       op.is_max, st);
    } else {
      ConvDesc d;
      fill_conv_desc(h, op, batch, input, outputs, d);
      const int lay = op.in.buf == -1 ? (in_layout == YOLO_IN_NCHW_F32 ? 1 : 2) : 0;
      if (op.umma.enabled) rc = launch_conv_umma(op.umma, d, st);
      else if (stem_eligible(d, lay)) rc = launch_stem(d, lay, st);
      else rc = launch_conv_simt(d, lay, st);
    }
    if (rc) return hfail(h, rc);
  }
  h->last_launches = g_launches - launches0;
  return YOLO_OK;
}

extern "C" int yolo_debug_activation(yolo_handle* h, const char* layer_name, int batch, float* host_nchw, size_t n_elems) {
  if (!h || !layer_name || !host_nchw) return fail(YOLO_E_BADARG, "debug_activation: null argument");
  if (strncmp(layer_name, "grad:", 5) == 0) return yb::train_debug_grad(h, layer_name + 5, batch, host_nchw, n_elems);
  auto it = h->named.find(layer_name);
  if (it == h->named.end()) return hfail(h, fail(YOLO_E_BADARG, "debug_activation: unknown layer '%s'", layer_name));
  const View& v = it->second;
  if (v.buf < 0) return hfail(h, fail(YOLO_E_BADARG, "debug_activation: '%s' is a user output", layer_name));
  if (!h->ws) return hfail(h, fail(YOLO_E_STATE, "debug_activation: no workspace"));
  if (n_elems != (size_t)batch * v.C * v.H * v.W) return hfail(h, fail(YOLO_E_SHAPE, "debug_activation: '%s' is (%d,%d,%d,%d)", layer_name, batch, v.C, v.H, v.W));
  YB_CUDA(cudaSetDevice(h->device));
  YB_CUDA(cudaDeviceSynchronize());
  const size_t esz = v.dtype == DT_F32 ? 4 : 2;
  const int nplanes = dtype_planes(v.dtype);
  const size_t nraw = (size_t)(nplanes - 1) * v.ps + ((size_t)batch * v.H * v.W - 1) * v.cpitch + v.coff + v.C;
  std::vector<unsigned char> raw(nraw * esz);
  YB_CUDA(cudaMemcpy(raw.data(), h->ws + h->bufs[v.buf].offset, raw.size(), cudaMemcpyDeviceToHost));
  for (int n = 0; n < batch; ++n)
    for (int y = 0; y < v.H; ++y)
      for (int x = 0; x < v.W; ++x)
        for (int c = 0; c < v.C; ++c) {
          size_t src = (((size_t)n * v.H + y) * v.W + x) * v.cpitch + v.coff + c;
          float f = 0.f;
          if (v.dtype != DT_F32) {
            for (int pl = 0; pl < nplanes; ++pl) {
              unsigned short u = reinterpret_cast<unsigned short*>(raw.data())[src + (size_t)pl * v.ps];
              float t;
              if (v.dtype == DT_F16X2) {
                __half hh;
                memcpy(&hh, &u, 2);
                t = __half2float(hh) * (pl == 1 ? kF16LoScaleInv : 1.f);
              } else {
                unsigned int w = (unsigned int)u << 16;
                memcpy(&t, &w, 4);
              }
              f += t;
            }
          } else f = reinterpret_cast<float*>(raw.data())[src];
          host_nchw[(((size_t)n * v.C + c) * v.H + y) * v.W + x] = f;
        }
  return YOLO_OK;
}

extern "C" int yolo_predict_host(yolo_handle* h, const void* host_input, int batch, int in_layout, float* host_rows,
                                 int32_t* host_idx, void* stream) {
  if (!h || !host_input || !host_rows) return fail(YOLO_E_BADARG, "predict_host: null argument");
  const yolo_spec& s = h->spec;
  if (s.net_type == YOLO_NET_LPDENSENET || s.net_type == YOLO_NET_CARDENSENET) return hfail(h, fail(YOLO_E_UNSUPPORTED, "predict_host: CARNET/CARLPNET only"));
  if (batch < 1 || batch > s.max_batch) return hfail(h, fail(YOLO_E_SHAPE, "predict_host: batch=%d outside [1,%d]", batch, s.max_batch));
  YB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t in_elem = in_layout == YOLO_IN_NCHW_F32 ? 4 : 1;
  const size_t in_bytes = (size_t)3 * s.height * s.width * in_elem;
  const int C = s.channels_per_anchor;
  // staging layout: input | outputs... | rows | idx   (all sized for max_batch, 256 B aligned)
  std::vector<size_t> off;
  size_t total = 0;
  auto take = [&](size_t b) { size_t o = total; total += (b + 255) & ~(size_t)255; return o; };
  const size_t o_in = take((size_t)3 * s.height * s.width * 4 * s.max_batch);
  std::vector<size_t> o_out(h->outputs.size());
  for (size_t i = 0; i < h->outputs.size(); ++i) {
    const View& v = h->outputs[i];
    o_out[i] = take((size_t)v.H * v.W * v.C * 4 * s.max_batch);
  }
  const size_t o_rows = take((size_t)C * 4 * s.max_batch), o_idx = take((size_t)4 * s.max_batch);
  if (!h->stage || h->stage_bytes < total) {
    if (h->stage) cudaFree(h->stage);
    h->stage = nullptr;
    if (cudaMalloc(&h->stage, total) != cudaSuccess) { cudaGetLastError(); return hfail(h, fail(YOLO_E_OOM, "predict_host: cudaMalloc(%zu) failed", total)); }
    h->stage_bytes = total;
  }
  char* sb = static_cast<char*>(h->stage);
  YB_CUDA(cudaMemcpyAsync(sb + o_in, host_input, in_bytes * batch, cudaMemcpyHostToDevice, st));
  std::vector<void*> outs(h->outputs.size());
  for (size_t i = 0; i < outs.size(); ++i) outs[i] = sb + o_out[i];
  int rc = yolo_forward(h, sb + o_in, batch, in_layout, outs.data(), stream);
  if (rc) return rc;
  yolo_decode_geom g;
  memset(&g, 0, sizeof(g));
  g.height = s.height; g.width = s.width; g.n_scales = s.n_scales; g.n_anchors = s.n_anchors; g.channels_per_anchor = C;
  for (int i = 0; i < s.n_scales; ++i) {
    g.step[i] = 1 << (s.n_layers - s.n_scales + 1 + i);       // car/YOLO.py:112-116
    for (int a = 0; a < s.n_anchors; ++a) { g.anchors[i][a][0] = s.anchors[i][a][0]; g.anchors[i][a][1] = s.anchors[i][a][1]; }
  }
  rc = yolo_decode_top1(&g, outs.data(), batch, reinterpret_cast<float*>(sb + o_rows), reinterpret_cast<int32_t*>(sb + o_idx), stream);
  if (rc) return hfail(h, rc);
  h->last_launches += 1;
  YB_CUDA(cudaMemcpyAsync(host_rows, sb + o_rows, (size_t)C * 4 * batch, cudaMemcpyDeviceToHost, st));
  if (host_idx) YB_CUDA(cudaMemcpyAsync(host_idx, sb + o_idx, (size_t)4 * batch, cudaMemcpyDeviceToHost, st));
  YB_CUDA(cudaStreamSynchronize(st));
  return YOLO_OK;
}

extern "C" int yolo_check_saturation(yolo_handle* h, int32_t* flags_out, void* stream) {
  if (!h || !flags_out) return fail(YOLO_E_BADARG, "check_saturation: null argument");
  *flags_out = 0;
  if (!h->d_flags) return YOLO_OK;
  YB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  int v = 0;
  YB_CUDA(cudaMemcpyAsync(&v, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  YB_CUDA(cudaMemsetAsync(h->d_flags, 0, sizeof(int), st));
  YB_CUDA(cudaStreamSynchronize(st));
  *flags_out = v;
  return YOLO_OK;
}

extern "C" int yolo_last_launch_count(const yolo_handle* h) { return h ? h->last_launches : YOLO_E_BADARG; }
extern "C" double yolo_conv_flops_per_image(const yolo_handle* h) { return h ? h->flops_per_image : 0.0; }
