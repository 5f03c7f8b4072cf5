// One data-parallel training step of the YOLO net on one GPU on the tensor cores.
//
// Replaces (reference, file:line): `_train_batch` car/YOLO.py:350-399 - forward with batch-statistics BatchNorm
// (per device, num_sync_bn_devices=-1 :94-96), `_loss_mask`/`_get_loss` (-> train_loss.cu), `sum(losses).backward()` :394
// (cuDNN bwd-data / bwd-filter / BN backward dispatched by MXNet autograd) and `trainer.step(batch_size)` :396 (kvstore
// reduce + `adam_update`).
//
// Arithmetic: fp32-grade throughout, on tcgen05.  Activations (y), raw convolution outputs (z) and the pre-activation
// gradients (dz) are stored as two fp16 planes v = hi + lo (the inference format, conv_umma.cu); activation gradients (dy)
// stay fp32 because several consumers accumulate into them.
//   forward   conv_umma_kernel writes z and, in its epilogue, the per-warp BatchNorm partial sums (shuffle butterfly); a
//             fixed-order reduction finalises mean / rstd / running statistics; one pass normalises + LeakyReLU (+ residual)
//             through the same channel-slice views (route concat, 2x upsample) as the inference plan.
//   backward  per layer, in reverse: BatchNorm backward = one reduction pass (sum g, sum g*xhat, max|g|; also routes the
//             residual's gradient) + one pass that writes dz as fp16 planes scaled by a per-layer power of two chosen ON THE
//             DEVICE (gradients span decades; the fp16 planes do not) - the scale is undone in the consumers' epilogues;
//             data gradient = conv_umma_kernel on the flipped filter, accumulating into the fp32 dy buffer (strided convs
//             read a zero-dilated copy of dz); weight gradient = wgrad_umma_kernel (MN-major operands, deterministic).
//   update    fused rescale + Adam on ONE flat parameter/gradient buffer, then the fp16 weight planes of both convolution
//             directions are re-packed on the device and the inference epilogues refolded: net.forward serves the trained net.
// Everything is deterministic (no floating-point atomics).  The few shapes the tensor-core kernels do not take (3-channel stem,
// Cin = 32, the 90-channel head convs: ~3 % of the FLOPs) run on the FFMA kernels.
//
// Data parallelism: the gradient is summed over ranks bucket by bucket WHILE the backward runs (yolo_train_comm_init): when
// the weight gradients of a bucket's layers are complete an event releases ncclAllReduce of that slice on a communication
// stream (NCCL over NVLink); the update waits for the last bucket.  This mirrors kvstore 'device' + trainer.step (car/YOLO.py:396).
#include <cuda_fp16.h>
#include <dlfcn.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "net_internal.cuh"
#include "wgrad_umma.cuh"

namespace yb {

constexpr float kBnMomentum = 0.9f;       // gluon BatchNorm() default (TrainState::bn_momentum)
constexpr float kTrainBnEps = 1e-5f;
constexpr int kGuardRows = 128;           // zero rows behind the last pixel of dz (tiles of the tensor-core kernels run past M)
constexpr int kSlabCtas = 592;             // CTAs a slab-reduction launch aims for (4 per SM)

// ---- minimal NCCL binding, resolved at run time from the libnccl the process already has (torch's) -------------------------
typedef struct ncclComm* ncclComm_t;
struct NcclUid { char b[128]; };           // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUid, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
  if (g_nccl.lib) return YOLO_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);         // prefer the copy already loaded (torch's bundled NCCL)
    if (!lib) lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return fail(YOLO_E_NCCL, "NCCL: libnccl.so.2 not found (%s)", dlerror());
  NcclApi a;
  a.lib = lib;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(lib, "ncclAllReduce"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy || !a.GetErrorString)
    return fail(YOLO_E_NCCL, "NCCL: required symbols missing in libnccl");
  g_nccl = a;
  return YOLO_OK;
}
#define YB_NCCL(expr)                                                                                          \
  do {                                                                                                         \
    int _r = (expr);                                                                                           \
    if (_r != 0) return yb::fail(YOLO_E_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
  } while (0)

struct TrainLayer {                       // per conv op
  // offsets (in floats) into the flat parameter / gradient buffers
  size_t o_w = 0, o_gamma = 0, o_beta = 0, o_bias = 0;
  bool has_bn = false, has_bias = false;
  int Ho = 0, Wo = 0;
  // forward
  __half* z = nullptr;  long long z_ps = 0;          // raw conv output, fp16 planes [2][max_batch*Ho*Wo][Cout]
  float* zf = nullptr;                                // head convs: fp32 output (= the head tensor)
  float* stat_part = nullptr;  size_t stat_groups = 0;   // per-warp partial sums [groups][2][Cout]
  double* slab = nullptr;  int slab_cap = 1;          // [slab_cap][2][Cout] second-level partials (forward and backward)
  float* mean = nullptr; float* rstd = nullptr; float* rmean = nullptr; float* rvar = nullptr;
  // backward
  float* ab = nullptr;                                // [2][Cout]: a = gamma*rstd, b = beta - mean*a  (u = a*z + b)
  float* mg = nullptr;                                // [2][Cout]: mean g, mean g*xhat
  float* dzscale = nullptr;                           // device: [0] = 2^s applied to dz, [1] = 2^-s, [2] = max|g| bits, [3] = max|gamma*rstd| bits (uint)
  __half* dz = nullptr;  long long dz_plane_rows = 0; // scaled pre-activation gradient [2][dz_plane_rows][Cout]
  __half* dzd = nullptr; long long dzd_plane_rows = 0; int dzd_n = 0;   // stride > 1: the same, zero-dilated to the input resolution
  UmmaConv dgrad;                                     // data-gradient convolution (Cout -> Cin, flipped filter)
  UmmaConv dgrad_par[4];                              // 3x3 / stride 2 / pad 1: one dense convolution over dz per input-pixel parity class
  bool dgrad_parity = false;
  WgradPlan wg;
  float* wT = nullptr; int cin_pad = 0;               // FFMA data gradient (head convs): flipped fp32 weights
  bool fwd_umma = false;
  bool dgrad_accum = true, dres_accum = true;         // false: this op is the FIRST writer of that gradient range -> plain store, no zero-fill needed
};

struct TrainState {
  float* P = nullptr; float* G = nullptr; float* M1 = nullptr; float* M2 = nullptr;    // caller-owned flat buffers
  size_t n_flat = 0;
  std::vector<TrainLayer> layers;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<size_t> gbuf_off;           // byte offset of the fp32 gradient buffer mirroring forward buffer i
  size_t grads_begin = 0, grads_bytes = 0;
  std::vector<char> gbuf_memset;          // gradient buffers that still need zero-filling each step (first-writer analysis failed)
  void* loss_scratch = nullptr; size_t loss_scratch_bytes = 0;
  float* dheads[YOLO_MAX_SCALES + 1] = {nullptr, nullptr, nullptr, nullptr};
  float* wg_scratch = nullptr; size_t wg_scratch_bytes = 0;
  int step_count = 0;
  // gradient all-reduce
  ncclComm_t comm = nullptr; int world = 1, rank = 0;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_bucket = nullptr, ev_done = nullptr;
  size_t bucket_floats = 16u << 20;       // 64 MB buckets
  bool reduced = false;
  float bn_momentum = kBnMomentum;
};

// ---------------------------------------------------------------------------------------------------
// device helpers: eight consecutive channels of an fp16x2 tensor
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld8_f16x2(const __half* p, long long ps, float v[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p + ps));
  const __half2* ha = reinterpret_cast<const __half2*>(&a);
  const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = __half22float2(ha[i]), y = __half22float2(hb[i]);
    v[2 * i] = x.x + y.x; v[2 * i + 1] = x.y + y.y;
  }
}
__device__ __forceinline__ void st8_f16x2(__half* p, long long ps, const float v[8], int& sat) {
  uint4 a, b;
  __half2* ha = reinterpret_cast<__half2*>(&a);
  __half2* hb = reinterpret_cast<__half2*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float x = v[2 * i], y = v[2 * i + 1];
    if (fmaxf(fabsf(x), fabsf(y)) > kF16Max) sat = 1;
    const __half2 h = __floats2half2_rn(fminf(fmaxf(x, -kF16Max), kF16Max), fminf(fmaxf(y, -kF16Max), kF16Max));
    const float2 f = __half22float2(h);
    ha[i] = h;
    hb[i] = __floats2half2_rn(x - f.x, y - f.y);
  }
  *reinterpret_cast<uint4*>(p) = a;
  *reinterpret_cast<uint4*>(p + ps) = b;
}
__device__ __forceinline__ void ld8_f32(const float* p, float v[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ float act_grad(float g, float u, int act) {
  if (act == ACT_LEAKY) return u > 0.f ? g : 0.1f * g;
  if (act == ACT_RELU) return u > 0.f ? g : 0.f;
  return g;
}

// ---------------------------------------------------------------------------------------------------
// forward kernels
// ---------------------------------------------------------------------------------------------------
// Per-warp partial sums of z (layers whose convolution ran on the FFMA kernels): one warp = 32 rows x 32 channels,
// written in the layout of conv_umma_kernel's epilogue statistics: part[group][0 | 1][C].
__global__ void __launch_bounds__(256)
bn_stats_rows_kernel(const __half* __restrict__ z, long long z_ps, int M, int C, float* __restrict__ part) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int cchunks = (C + 31) >> 5;
  const int grp = warp / cchunks, c = (warp - grp * cchunks) * 32 + lane;
  if ((long long)grp * 32 >= M) return;
  float s1 = 0.f, s2 = 0.f;
  if (c < C) {
    const int m1 = min(M, grp * 32 + 32);
    for (int m = grp * 32; m < m1; ++m) {
      const size_t i = (size_t)m * C + c;
      const float v = __half2float(z[i]) + __half2float(z[z_ps + i]);
      s1 += v; s2 += v * v;
    }
    part[((size_t)grp * 2) * C + c] = s1;
    part[((size_t)grp * 2 + 1) * C + c] = s2;
  }
}

// second level: slab y of the row groups -> double partials slab[y][q][C]; fixed order, no atomics
__global__ void __launch_bounds__(256)
reduce_groups_kernel(const float* __restrict__ part, size_t groups, int C, double* __restrict__ slab) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  const size_t per = (groups + gridDim.y - 1) / gridDim.y;
  const size_t g0 = blockIdx.y * per, g1 = min(groups, g0 + per);
  double s0 = 0.0, s1 = 0.0;
  if (c < C)
    for (size_t g = g0 + rl; g < g1; g += 8) { s0 += part[(g * 2) * C + c]; s1 += part[(g * 2 + 1) * C + c]; }
  __shared__ double sh[2][8][32];
  sh[0][rl][threadIdx.x & 31] = s0; sh[1][rl][threadIdx.x & 31] = s1;
  __syncthreads();
  if (rl == 0 && c < C) {
    for (int r = 1; r < 8; ++r) { s0 += sh[0][r][threadIdx.x & 31]; s1 += sh[1][r][threadIdx.x & 31]; }
    slab[((size_t)blockIdx.y * 2) * C + c] = s0;
    slab[((size_t)blockIdx.y * 2 + 1) * C + c] = s1;
  }
}

// mean / rstd / running statistics / folded (a, b) from the slab partials: one warp per channel, lanes stride over the slabs,
// shuffle tree (fixed order -> deterministic)
__global__ void __launch_bounds__(256)
bn_finalize_fwd_kernel(const double* __restrict__ slab, int nslab, int M, int C, float* mean, float* rstd, float* rmean, float* rvar,
                       float momentum, const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ ab) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int y = lane; y < nslab; y += 32) { s0 += slab[((size_t)y * 2) * C + c]; s1 += slab[((size_t)y * 2 + 1) * C + c]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xFFFFFFFFu, s0, o); s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o); }
  if (lane) return;
  const double mu = s0 / M, var = fmax(s1 / M - mu * mu, 0.0);
  mean[c] = (float)mu;
  rstd[c] = (float)(1.0 / sqrt(var + (double)kTrainBnEps));
  rmean[c] = rmean[c] * momentum + (float)mu * (1.f - momentum);
  rvar[c] = rvar[c] * momentum + (float)var * (1.f - momentum);
  const float a = gamma[c] * rstd[c];                        // u = gamma*(z - mean)*rstd + beta = a*z + b
  ab[c] = a;
  ab[C + c] = beta[c] - mean[c] * a;
}

// y = act(a*z + b) (+ residual), written through the op's output view (concat slice / 2x upsample).  Block = (256 / octs) row lanes x
// octs channel octets over <= 256 channels (per-channel constants live in registers); grid = (channel blocks, row blocks).
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const __half* __restrict__ z, long long z_ps, int M, int C, const float* __restrict__ ab, int act,
                  const __half* __restrict__ res, int res_cpitch, int res_coff, long long res_ps, __half* __restrict__ out, int out_cpitch,
                  int out_coff, long long out_ps, int upsample2, int Ho, int Wo, int* sat_flag) {
  const int cb = min(C, 256), octs = cb >> 3, lanes = 256 / octs;
  const int oc = threadIdx.x % octs, rl = threadIdx.x / octs;
  const int c = blockIdx.x * cb + oc * 8;
  if (c >= C || rl >= lanes) return;
  float a[8], b[8];
  ld8_f32(ab + c, a);
  ld8_f32(ab + C + c, b);
  int sat = 0;
  for (int m = blockIdx.y * lanes + rl; m < M; m += gridDim.y * lanes) {
    float v[8], r[8];
    ld8_f16x2(z + (size_t)m * C + c, z_ps, v);
    if (res) ld8_f16x2(res + (size_t)m * res_cpitch + res_coff + c, res_ps, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float u = fmaf(v[j], a[j], b[j]);
      if (act == ACT_LEAKY) u = u > 0.f ? u : 0.1f * u;
      else if (act == ACT_RELU) u = fmaxf(u, 0.f);
      v[j] = res ? u + r[j] : u;
    }
    if (upsample2) {
      const int HoWo = Ho * Wo, n = m / HoWo, rem = m - n * HoWo, oh = rem / Wo, ow = rem - oh * Wo;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const size_t pix = ((size_t)n * 2 * Ho + 2 * oh + (q >> 1)) * (2 * Wo) + 2 * ow + (q & 1);
        st8_f16x2(out + pix * out_cpitch + out_coff + c, out_ps, v, sat);
      }
    } else st8_f16x2(out + (size_t)m * out_cpitch + out_coff + c, out_ps, v, sat);
  }
  if (sat && sat_flag) atomicOr(sat_flag, YOLO_SAT_ACT_BN);
}

// ---------------------------------------------------------------------------------------------------
// backward kernels
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_dy8(const float* dy, int dy_cpitch, int dy_coff, int upsample2, int Ho, int Wo, int m, int c, float g[8]) {
  if (upsample2) {
    const int HoWo = Ho * Wo, n = m / HoWo, rem = m - n * HoWo, oh = rem / Wo, ow = rem - oh * Wo;
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const size_t pix = ((size_t)n * 2 * Ho + 2 * oh + (q >> 1)) * (2 * Wo) + 2 * ow + (q & 1);
      float t[8];
      ld8_f32(dy + pix * dy_cpitch + dy_coff + c, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += t[j];
    }
  } else ld8_f32(dy + (size_t)m * dy_cpitch + dy_coff + c, g);
}

// Reduction pass of the BatchNorm backward: per channel sum g and sum g*z with g = dy * act'(u) (fixed-order slabs -> slab[y][q][C]; the
// finalize kernel centres the second moment in double: sum g*xhat = rstd * (sum g*z - mean * sum g)), the layer's max|g| (for the dz
// scale) and the residual's gradient (y = act(bn(z)) + res  ->  d res += dy).
// Block = (256 / octs) row lanes x octs channel octets over <= 256 channels; grid = (channel blocks, slabs).  A thread sums at most a
// few hundred rows (16-row fp32 partials, then fp32 totals); everything across threads and slabs is added in double.  Raw moments and
// fp32 thread totals keep the kernel at 3 blocks per SM (it ran at 128 registers and 3.3 TB/s with per-element xhat and double totals).
__global__ void __launch_bounds__(256, 3)
bn_bwd_reduce_kernel(const __half* __restrict__ z, long long z_ps, int M, int C, const float* __restrict__ dy, int dy_cpitch, int dy_coff,
                     int upsample2, int Ho, int Wo, const float* __restrict__ ab, int act, float* __restrict__ dres, int dres_cpitch,
                     int dres_coff, int dres_accum, double* __restrict__ slab, unsigned int* __restrict__ gmax_bits) {
  const int cb = min(C, 256), octs = cb >> 3, lanes = 256 / octs;
  const int oc = threadIdx.x % octs, rl = threadIdx.x / octs;
  const int c = blockIdx.x * cb + oc * 8;
  const int per = (M + gridDim.y - 1) / gridDim.y;
  const int m0 = blockIdx.y * per, m1 = min(M, m0 + per);
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s0[j] = 0.f; s1[j] = 0.f; }
  float gm = 0.f;
  if (c < C && rl < lanes) {
    float ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ga[j] = ab[c + j]; be[j] = ab[C + c + j]; }
    for (int mb = m0 + rl; mb < m1; mb += lanes * 16) {
      float f0[8], f1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { f0[j] = 0.f; f1[j] = 0.f; }
      for (int t = 0; t < 16; ++t) {
        const int m = mb + t * lanes;
        if (m >= m1) break;
        float zv[8], g[8];
        ld8_f16x2(z + (size_t)m * C + c, z_ps, zv);
        load_dy8(dy, dy_cpitch, dy_coff, upsample2, Ho, Wo, m, c, g);
        if (dres) {
          float* rp = dres + (size_t)m * dres_cpitch + dres_coff + c;
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
          if (dres_accum) { a = *reinterpret_cast<float4*>(rp); b = *reinterpret_cast<float4*>(rp + 4); }
          a.x += g[0]; a.y += g[1]; a.z += g[2]; a.w += g[3]; b.x += g[4]; b.y += g[5]; b.z += g[6]; b.w += g[7];
          *reinterpret_cast<float4*>(rp) = a; *reinterpret_cast<float4*>(rp + 4) = b;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gg = act_grad(g[j], fmaf(zv[j], ga[j], be[j]), act);        // u = a*z + b exactly as the forward computed it
          f0[j] += gg; f1[j] = fmaf(gg, zv[j], f1[j]);
          gm = fmaxf(gm, fabsf(gg));
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { s0[j] += f0[j]; s1[j] += f1[j]; }
    }
  }
  // row lanes -> one value per channel (shared memory; binary tree when the lane count is a power of two - fixed order either way)
  __shared__ double sh[256][9];
  const bool pow2 = (lanes & (lanes - 1)) == 0;
  for (int q = 0; q < 2; ++q) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[threadIdx.x][j] = (double)(q == 0 ? s0[j] : s1[j]);
    __syncthreads();
    if (pow2) {
      for (int st = lanes >> 1; st >= 1; st >>= 1) {
        if (rl < st) {
#pragma unroll
          for (int j = 0; j < 8; ++j) sh[threadIdx.x][j] += sh[(rl + st) * octs + oc][j];
        }
        __syncthreads();
      }
      if (rl == 0 && c < C) {
#pragma unroll
        for (int j = 0; j < 8; ++j) slab[((size_t)blockIdx.y * 2 + q) * C + c + j] = sh[threadIdx.x][j];
      }
    } else if (rl == 0 && c < C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double a = 0.0;
        for (int r = 0; r < lanes; ++r) a += sh[r * octs + oc][j];
        slab[((size_t)blockIdx.y * 2 + q) * C + c + j] = a;
      }
    }
  }
  // max|g|: order independent
  gm = fmaxf(gm, __shfl_xor_sync(0xFFFFFFFFu, gm, 16));
  gm = fmaxf(gm, __shfl_xor_sync(0xFFFFFFFFu, gm, 8));
  gm = fmaxf(gm, __shfl_xor_sync(0xFFFFFFFFu, gm, 4));
  gm = fmaxf(gm, __shfl_xor_sync(0xFFFFFFFFu, gm, 2));
  gm = fmaxf(gm, __shfl_xor_sync(0xFFFFFFFFu, gm, 1));
  if ((threadIdx.x & 31) == 0 && gm > 0.f) atomicMax(gmax_bits, __float_as_uint(gm));
}

// per channel: mean g, mean g*xhat, dgamma, dbeta (one warp per channel, fixed-order shuffle tree) and the layer's max|gamma*rstd|
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const double* __restrict__ slab, int nslab, int M, int C, const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ ab, float* __restrict__ mg, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       unsigned int* __restrict__ amax_bits) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int y = lane; y < nslab; y += 32) { s0 += slab[((size_t)y * 2) * C + c]; s1 += slab[((size_t)y * 2 + 1) * C + c]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xFFFFFFFFu, s0, o); s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o); }
  if (lane) return;
  s1 = (double)rstd[c] * (s1 - (double)mean[c] * s0);         // sum g*xhat from the raw moments
  mg[c] = (float)(s0 / M); mg[C + c] = (float)(s1 / M);
  dbeta[c] = (float)s0; dgamma[c] = (float)s1;
  atomicMax(amax_bits, __float_as_uint(fabsf(ab[c])));        // order independent
}

// power-of-two dz scale from the bound |dz| <= max_c|gamma*rstd| * max|g| * 8 (|xhat| <= 6 assumed; the fp16 high plane still has
// 16x headroom above the bound, and saturation is flagged)
__device__ __forceinline__ float dz_scale_from(const float* dzscale) {
  const float gmax = __uint_as_float(reinterpret_cast<const unsigned int*>(dzscale)[2]);
  const float amax = __uint_as_float(reinterpret_cast<const unsigned int*>(dzscale)[3]);
  const float bound = amax * gmax * 8.f;
  float s = 1.f;
  if (bound > 0.f && isfinite(bound)) {
    int e;
    frexpf(bound, &e);                                      // bound = f * 2^e, f in [0.5, 1)
    int sh = 12 - e;
    sh = sh < -100 ? -100 : (sh > 100 ? 100 : sh);
    s = ldexpf(1.f, sh);
  }
  return s;
}

// dz = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)) * 2^s as fp16 planes (+ the zero-dilated copy for strided convolutions);
// also zeroes the guard rows [M, M + kGuardRows) of dz.  Thread mapping as in bn_act_fwd_kernel.
__global__ void __launch_bounds__(256, 3)
bn_bwd_apply_kernel(const __half* __restrict__ z, long long z_ps, int M, int C, const float* __restrict__ mg, const float* __restrict__ dy,
                    int dy_cpitch, int dy_coff, int upsample2, int Ho, int Wo, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ ab, int act, float* __restrict__ dzscale,
                    __half* __restrict__ dz, long long dz_ps, __half* __restrict__ dzd, long long dzd_ps, int stride, int Hin, int Win,
                    int* sat_flag) {
  const int cb = min(C, 256), octs = cb >> 3, lanes = 256 / octs;
  const int oc = threadIdx.x % octs, rl = threadIdx.x / octs;
  const int c = blockIdx.x * cb + oc * 8;
  if (c >= C || rl >= lanes) return;
  const float s = dz_scale_from(dzscale);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { dzscale[0] = s; dzscale[1] = 1.f / s; }   // read by the dgrad / wgrad epilogues
  // dz*s = a*s*(gg - m0 - xhat*m1) = k1*gg + k2*z + k3 with per-channel constants (three registers per channel instead of six)
  float a[8], b[8], k1[8], k2[8], k3[8];
  {
    float mu[8], rs[8], m0[8], m1[8];
    ld8_f32(mean + c, mu); ld8_f32(rstd + c, rs); ld8_f32(ab + c, a); ld8_f32(ab + C + c, b); ld8_f32(mg + c, m0); ld8_f32(mg + C + c, m1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      k1[j] = a[j] * s;                                               // a = gamma * rstd
      k2[j] = -k1[j] * m1[j] * rs[j];
      k3[j] = k1[j] * (m1[j] * rs[j] * mu[j] - m0[j]);
    }
  }
  int sat = 0;
  for (int m = blockIdx.y * lanes + rl; m < M + kGuardRows; m += gridDim.y * lanes) {
    float o[8];
    if (m < M) {
      float zv[8], g[8];
      ld8_f16x2(z + (size_t)m * C + c, z_ps, zv);
      load_dy8(dy, dy_cpitch, dy_coff, upsample2, Ho, Wo, m, c, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gg = act_grad(g[j], fmaf(zv[j], a[j], b[j]), act);
        o[j] = fmaf(k1[j], gg, fmaf(k2[j], zv[j], k3[j]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    }
    st8_f16x2(dz + (size_t)m * C + c, dz_ps, o, sat);
    if (dzd && m < M) {
      const int HoWo = Ho * Wo, n = m / HoWo, rem = m - n * HoWo, oh = rem / Wo, ow = rem - oh * Wo;
      const size_t pix = ((size_t)n * Hin + (size_t)oh * stride) * Win + (size_t)ow * stride;
      st8_f16x2(dzd + pix * C + c, dzd_ps, o, sat);
    }
  }
  if (sat && sat_flag) atomicOr(sat_flag, YOLO_SAT_GRAD);
}

// column sums of a dense fp32 [M][C] matrix (bias gradient of the head convs): slab partials (double) + a warp-per-channel finalize
__global__ void __launch_bounds__(256)
colsum_slab_kernel(const float* __restrict__ a, int M, int C, double* __restrict__ slab) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  const int per = (M + gridDim.y - 1) / gridDim.y;
  const int m0 = blockIdx.y * per, m1 = min(M, m0 + per);
  double s = 0.0;
  if (c < C)
    for (int m = m0 + rl; m < m1; m += 8) s += a[(size_t)m * C + c];
  __shared__ double sh[8][32];
  sh[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < C) {
    for (int r = 1; r < 8; ++r) s += sh[r][threadIdx.x & 31];
    slab[(size_t)blockIdx.y * C + c] = s;
  }
}
__global__ void __launch_bounds__(256)
colsum_finalize_kernel(const double* __restrict__ slab, int nslab, int C, float* __restrict__ out) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0;
  for (int y = lane; y < nslab; y += 32) s += slab[(size_t)y * C + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if (lane == 0) out[c] = (float)s;
}

// FFMA weight gradient for the shapes the tensor-core kernel does not take: dW[k][n] = sum_m A[m][k] * dz[m][n], A = im2col gather of
// the layer input (fp32 NCHW / uint8 NHWC network input, or fp16 planes); dz fp32 dense or scaled fp16 planes.  64 x 64 output tile
// per CTA, split over M into gridDim.z slabs that write partial tiles (summed in a fixed order by wgrad_reduce).
template <bool XF16, bool DF16>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const void* __restrict__ xv, long long x_ps, int N, int H, int W, int Cin, int x_cpitch, int x_coff, int in_layout,
                  const void* __restrict__ dzv, long long dz_ps, int dz_cpitch, int Ho, int Wo, int Cout, int kh, int kw, int stride, int pad,
                  float* __restrict__ part, size_t slab_stride, int cout_pad) {
  constexpr int TK = 64, TN = 64, TM = 16;
  __shared__ float As[TM][TK + 1];
  __shared__ float Ds[TM][TN];
  const int K = kh * kw * Cin, M = N * Ho * Wo, HoWo = Ho * Wo;
  const int k0 = blockIdx.x * TK, n0 = blockIdx.y * TN;
  const int slabM = (M + gridDim.z - 1) / gridDim.z;
  const int m_begin = blockIdx.z * slabM, m_end = min(M, m_begin + slabM);
  const int tid = threadIdx.x, tk = tid >> 4, tn = tid & 15;        // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int mb = m_begin; mb < m_end; mb += TM) {
    for (int e = tid; e < TM * TK; e += 256) {                      // gather A chunk
      const int ml = e / TK, kl = e - ml * TK;
      const int m = mb + ml, k = k0 + kl;
      float v = 0.f;
      if (m < m_end && k < K) {
        const int tap = k / Cin, c = k - tap * Cin, r = tap / kw, s = tap - r * kw;
        const int n = m / HoWo, rem = m - n * HoWo, oh = rem / Wo, ow = rem - oh * Wo;
        const int ih = oh * stride - pad + r, iw = ow * stride - pad + s;
        if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) {
          if (in_layout == 1) v = static_cast<const float*>(xv)[((size_t)(n * Cin + c) * H + ih) * W + iw];                        // network input NCHW fp32
          else if (in_layout == 2) v = (float)static_cast<const unsigned char*>(xv)[((size_t)(n * H + ih) * W + iw) * Cin + c] / 255.f;
          else if (XF16) {
            const __half* xp = static_cast<const __half*>(xv) + ((size_t)(n * H + ih) * W + iw) * x_cpitch + x_coff + c;
            v = __half2float(xp[0]) + __half2float(xp[x_ps]);
          } else v = static_cast<const float*>(xv)[((size_t)(n * H + ih) * W + iw) * x_cpitch + x_coff + c];
        }
      }
      As[ml][kl] = v;
    }
    for (int e = tid; e < TM * TN; e += 256) {
      const int ml = e / TN, nl = e - ml * TN;
      const int m = mb + ml, n = n0 + nl;
      float v = 0.f;
      if (m < m_end && n < Cout) {
        if (DF16) {
          const __half* dp = static_cast<const __half*>(dzv) + (size_t)m * dz_cpitch + n;
          v = __half2float(dp[0]) + __half2float(dp[dz_ps]);
        } else v = static_cast<const float*>(dzv)[(size_t)m * dz_cpitch + n];
      }
      Ds[ml][nl] = v;
    }
    __syncthreads();
#pragma unroll
    for (int ml = 0; ml < TM; ++ml) {
      float a[4], d[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[ml][tk * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = Ds[ml][tn * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], d[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* op = part + (size_t)blockIdx.z * slab_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + tk * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn * 4 + j;
      if (n < cout_pad) op[(size_t)k * cout_pad + n] = n < Cout ? acc[i][j] : 0.f;
    }
  }
}
// Weight gradient of the 3-channel stem (K = kh*kw*3 <= 32 rows, Cout <= 32): the 64 x 64 tile of wgrad_simt_kernel would be 80 % empty.
// Each block owns a slab of pixels; per chunk of 64 pixels the im2col patch (coalesced along the image row) and the dz rows are staged
// in shared memory and every thread accumulates its (k, n) outputs.  Slab partials are summed in a fixed order by wgrad_reduce_scaled.
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const void* __restrict__ xv, int in_layout, int N, int H, int W, const __half* __restrict__ dz, long long dz_ps, int Ho, int Wo,
                  int Cout, int kh, int kw, int stride, int pad, float* __restrict__ part, size_t slab_stride, int cout_pad) {
  constexpr int TM = 64, KMAX = 32, NMAX = 32;
  __shared__ float xs[KMAX][TM + 1];
  __shared__ float ds[TM][NMAX + 1];
  const int K = kh * kw * 3, M = N * Ho * Wo, HoWo = Ho * Wo;
  const int slabM = (M + gridDim.x - 1) / gridDim.x;
  const int m_begin = blockIdx.x * slabM, m_end = min(M, m_begin + slabM);
  const int tid = threadIdx.x;
  const int npairs = K * Cout;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int pk[4], pn[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { const int idx = tid + 256 * i; pk[i] = idx < npairs ? idx / Cout : 0; pn[i] = idx < npairs ? idx % Cout : 0; }
  for (int mb = m_begin; mb < m_end; mb += TM) {
    for (int e = tid; e < K * TM; e += 256) {
      const int k = e / TM, ml = e - k * TM, m = mb + ml;
      float v = 0.f;
      if (m < m_end) {
        const int tap = k / 3, c = k - tap * 3, r = tap / kw, sx = tap - r * kw;
        const int n = m / HoWo, rem = m - n * HoWo, oh = rem / Wo, ow = rem - oh * Wo;
        const int ih = oh * stride - pad + r, iw = ow * stride - pad + sx;
        if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) {
          if (in_layout == 1) v = __ldg(static_cast<const float*>(xv) + ((size_t)(n * 3 + c) * H + ih) * W + iw);
          else v = (float)__ldg(static_cast<const unsigned char*>(xv) + ((size_t)(n * H + ih) * W + iw) * 3 + c) / 255.f;
        }
      }
      xs[k][ml] = v;
    }
    for (int e = tid; e < TM * Cout; e += 256) {
      const int ml = e / Cout, n = e - ml * Cout, m = mb + ml;
      float v = 0.f;
      if (m < m_end) { const __half* dp = dz + (size_t)m * Cout + n; v = __half2float(dp[0]) + __half2float(dp[dz_ps]); }
      ds[ml][n] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int ml = 0; ml < TM; ++ml) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(xs[pk[i]][ml], ds[ml][pn[i]], acc[i]);
    }
    __syncthreads();
  }
  float* op = part + (size_t)blockIdx.x * slab_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + 256 * i;
    if (idx < npairs) op[(size_t)pk[i] * cout_pad + pn[i]] = acc[i];
  }
}
__global__ void wgrad_reduce_scaled_kernel(const float* __restrict__ part, int slabs, size_t n, size_t slab_stride, const float* __restrict__ scale_dev,
                                           float* __restrict__ out) {
  const float sc = scale_dev ? __ldg(scale_dev) : 1.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float a = part[i];
    for (int s = 1; s < slabs; ++s) a += part[(size_t)s * slab_stride + i];
    out[i] = a * sc;
  }
}

// wT[((kh-1-r)*kw + (kw-1-s))*Cout + o][c] = W[(r*kw+s)*Cin + c][o]
__global__ void wflip_kernel(const float* __restrict__ Wm, int kh, int kw, int Cin, int Cout, int cout_pad, float* __restrict__ wT, int cin_pad) {
  const size_t total = (size_t)kh * kw * Cin * Cout;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % Cout);
    size_t t = i / Cout;
    const int c = (int)(t % Cin);
    const int tap = (int)(t / Cin), r = tap / kw, s = tap - r * kw;
    wT[((size_t)((kh - 1 - r) * kw + (kw - 1 - s)) * Cout + o) * cin_pad + c] = Wm[((size_t)tap * Cin + c) * cout_pad + o];
  }
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m1, float& m2, float lr_t, float rescale, float b1, float b2, float eps) {
  g *= rescale;
  const float m = b1 * m1 + (1.f - b1) * g;
  const float v = b2 * m2 + (1.f - b2) * g * g;
  m1 = m; m2 = v;
  p -= lr_t * m / (sqrtf(v) + eps);
}
// 28 bytes of traffic per parameter: four floats per thread keep enough bytes in flight (n is a multiple of 4, the arrays 16-byte aligned)
__global__ void __launch_bounds__(256)
adam_kernel(float4* __restrict__ P, const float4* __restrict__ G, float4* __restrict__ M1, float4* __restrict__ M2, size_t n4,
            float lr_t, float rescale, float b1, float b2, float eps) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 p = P[i], m1 = M1[i], m2 = M2[i];
    const float4 g = G[i];
    adam_one(p.x, g.x, m1.x, m2.x, lr_t, rescale, b1, b2, eps);
    adam_one(p.y, g.y, m1.y, m2.y, lr_t, rescale, b1, b2, eps);
    adam_one(p.z, g.z, m1.z, m2.z, lr_t, rescale, b1, b2, eps);
    adam_one(p.w, g.w, m1.w, m2.w, lr_t, rescale, b1, b2, eps);
    M1[i] = m1; M2[i] = m2; P[i] = p;
  }
}

// inference scale/shift of the conv epilogues from the trained BN parameters and running statistics
__global__ void refold_kernel(const float* gamma, const float* beta, const float* rmean, const float* rvar, const float* bias, int C,
                              float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (gamma) {
    const float s = gamma[c] / sqrtf(rvar[c] + kTrainBnEps);
    scale[c] = s;
    shift[c] = beta[c] - rmean[c] * s + (bias ? bias[c] * s : 0.f);
  } else {
    scale[c] = 1.f;
    shift[c] = bias ? bias[c] : 0.f;
  }
}

static inline int grid_for(size_t n, int block = 256, int cap = 148 * 8) {
  size_t g = (n + block - 1) / block;
  return (int)(g < (size_t)cap ? (g ? g : 1) : cap);
}

void train_release(yolo_handle* h, bool writeback) {
  if (!h || !h->train) return;
  TrainState* T = h->train;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  // inference keeps working on the handle's own copies: restore the FFMA weight pointers into the parameter arena and refresh them
  // (and the host copies that yolo_get_param serves) from the trained flat buffer
  for (size_t i = 0; i < h->ops.size(); ++i) {
    Op& op = h->ops[i];
    if (op.kind != OP_CONV) continue;
    TrainLayer& L = T->layers[i];
    const size_t K = (size_t)op.kh * op.kw * op.in.C;
    umma_release(L.dgrad);
    for (UmmaConv& u : L.dgrad_par) umma_release(u);
    if (!writeback) { if (op.w_f32_own) op.w_f32 = op.w_f32_own; continue; }
    if (op.w_f32_own) {
      cudaMemcpy(op.w_f32_own, T->P + L.o_w, K * op.cout_pad * 4, cudaMemcpyDeviceToDevice);
      op.w_f32 = op.w_f32_own;
    }
    std::vector<float> tmp(K * op.cout_pad);
    if (cudaMemcpy(tmp.data(), T->P + L.o_w, tmp.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess) {
      if (!op.w_stem_host.empty() && op.w_stem_host.size() == tmp.size()) op.w_stem_host = tmp;      // kernel-parameter copy of the stem
      std::vector<float>& Wh = h->params[op.p_weight].host;
      for (int o = 0; o < op.cout; ++o)
        for (int c = 0; c < op.in.C; ++c)
          for (int r = 0; r < op.kh; ++r)
            for (int s2 = 0; s2 < op.kw; ++s2)
              Wh[(((size_t)o * op.in.C + c) * op.kh + r) * op.kw + s2] = tmp[((size_t)(r * op.kw + s2) * op.in.C + c) * op.cout_pad + o];
    }
    if (L.has_bn) {
      cudaMemcpy(h->params[op.p_bn].host.data(), T->P + L.o_gamma, op.cout * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(h->params[op.p_bn + 1].host.data(), T->P + L.o_beta, op.cout * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(h->params[op.p_bn + 2].host.data(), L.rmean, op.cout * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(h->params[op.p_bn + 3].host.data(), L.rvar, op.cout * 4, cudaMemcpyDeviceToHost);
    }
    if (L.has_bias) cudaMemcpy(h->params[op.p_bias].host.data(), T->P + L.o_bias, op.cout * 4, cudaMemcpyDeviceToHost);
  }
  cudaGetLastError();
  if (T->comm) g_nccl.CommDestroy(T->comm);
  if (T->comm_stream) cudaStreamDestroy(T->comm_stream);
  if (T->ev_bucket) cudaEventDestroy(T->ev_bucket);
  if (T->ev_done) cudaEventDestroy(T->ev_done);
  if (T->arena) cudaFree(T->arena);
  delete T;
  h->train = nullptr;
}

// slabs of the FFMA weight-gradient kernel: split the pixel range until ~4 CTAs per SM exist
static inline int simt_wgrad_slabs(int M, int K, int cout_pad) {
  const int tiles = ((K + 63) / 64) * ((cout_pad + 63) / 64);
  return std::max(1, std::min(M / 256, std::max(1, 592 / tiles)));
}
static inline int grad_pitch(const View& v) { return v.il ? v.cpitch / 2 : v.cpitch; }
static float* grad_ptr(const yolo_handle* h, const View& v) {
  return reinterpret_cast<float*>(h->train->arena + h->train->gbuf_off[v.buf]);
}
static __half* act16(const yolo_handle* h, const View& v) { return reinterpret_cast<__half*>(h->ws + h->bufs[v.buf].offset); }

int train_debug_grad(yolo_handle* h, const char* layer_name, int batch, float* host_nchw, size_t n_elems) {
  if (!h->train) return hfail(h, fail(YOLO_E_STATE, "debug_activation: no training state for '%s'", layer_name));
  auto it = h->named.find(layer_name);
  if (it == h->named.end()) return hfail(h, fail(YOLO_E_BADARG, "debug_activation: unknown layer '%s'", layer_name));
  const View& v = it->second;
  if (v.buf < 0) return hfail(h, fail(YOLO_E_BADARG, "debug_activation: '%s' is a user output", layer_name));
  if (n_elems != (size_t)batch * v.C * v.H * v.W) return hfail(h, fail(YOLO_E_SHAPE, "debug_activation: '%s' is (%d,%d,%d,%d)", layer_name, batch, v.C, v.H, v.W));
  YB_CUDA(cudaSetDevice(h->device));
  YB_CUDA(cudaDeviceSynchronize());
  const int pitch = grad_pitch(v);
  std::vector<float> raw((size_t)batch * v.H * v.W * pitch);
  YB_CUDA(cudaMemcpy(raw.data(), grad_ptr(h, v), raw.size() * 4, cudaMemcpyDeviceToHost));
  for (int n = 0; n < batch; ++n)
    for (int y = 0; y < v.H; ++y)
      for (int x = 0; x < v.W; ++x)
        for (int c = 0; c < v.C; ++c)
          host_nchw[(((size_t)n * v.C + c) * v.H + y) * v.W + x] = raw[(((size_t)n * v.H + y) * v.W + x) * pitch + v.coff + c];
  return YOLO_OK;
}

}  // namespace yb

using namespace yb;

extern "C" size_t yolo_train_flat_size(const yolo_handle* h) {
  if (!h) return 0;
  size_t n = 0;
  for (const Op& op : h->ops) {
    if (op.kind != OP_CONV) continue;
    n += (size_t)op.kh * op.kw * op.in.C * ((op.cout + 3) & ~3);
    if (op.p_bn >= 0) n += 2 * (size_t)op.cout;
    if (op.p_bias >= 0) n += op.cout;
    n = (n + 3) & ~(size_t)3;
  }
  return n;
}

// (re)pack the fp16 weight planes of both convolution directions from the flat fp32 parameters
static int repack_weights(yolo_handle* h, cudaStream_t st) {
  TrainState* T = h->train;
  for (size_t i = 0; i < h->ops.size(); ++i) {
    Op& op = h->ops[i];
    TrainLayer& L = T->layers[i];
    int rc = umma_pack_device(&op.umma, L.dgrad_parity ? L.dgrad_par : &L.dgrad, L.dgrad_parity ? 4 : 1, T->P + L.o_w, op.in.C, op.cout, op.cout_pad,
                              op.kh, op.kw, h->d_flags, st);
    if (rc) return rc;
    if (L.wT) {
      wflip_kernel<<<grid_for((size_t)op.kh * op.kw * op.in.C * op.cout), 256, 0, st>>>(T->P + L.o_w, op.kh, op.kw, op.in.C, op.cout, op.cout_pad, L.wT, L.cin_pad);
      ++g_launches;
    }
  }
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

extern "C" int yolo_train_init(yolo_handle* h, float* params_flat, float* grads_flat, float* adam_m, float* adam_v, size_t n_flat, void* stream) {
  if (!h || !params_flat || !grads_flat || !adam_m || !adam_v) return fail(YOLO_E_BADARG, "train_init: null argument");
  if (!h->finalized || !h->ws) return hfail(h, fail(YOLO_E_STATE, "train_init: parameters must be finalized and the workspace set"));
  if (h->spec.precision != YOLO_PREC_FP16X3) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train_init: training runs in YOLO_PREC_FP16X3 (fp32-grade on the tensor cores)"));
  if (h->spec.net_type != YOLO_NET_CARNET && h->spec.net_type != YOLO_NET_CARLPNET) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train_init: CARNET / CARLPNET only"));
  if (n_flat != yolo_train_flat_size(h)) return hfail(h, fail(YOLO_E_SHAPE, "train_init: flat buffers must hold %zu floats", yolo_train_flat_size(h)));
  if ((reinterpret_cast<uintptr_t>(params_flat) | reinterpret_cast<uintptr_t>(grads_flat) | reinterpret_cast<uintptr_t>(adam_m) |
       reinterpret_cast<uintptr_t>(adam_v)) & 15)
    return hfail(h, fail(YOLO_E_BADARG, "train_init: the flat buffers must be 16-byte aligned"));
  for (const Op& op : h->ops)
    if (op.kind != OP_CONV || op.p_prebn >= 0 || op.out_nchw || (op.p_bn >= 0 && op.cout % 8)) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train_init: unsupported op in the plan"));
  YB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  train_release(h, true);                   // re-initialisation continues from the trained parameters
  TrainState* T = new TrainState();
  h->train = T;
  T->P = params_flat; T->G = grads_flat; T->M1 = adam_m; T->M2 = adam_v; T->n_flat = n_flat;
  const int B = h->spec.max_batch;
  int num_sms = 0;
  int rc = device_sm_count(&num_sms);
  if (rc) return hfail(h, rc);
  // ---- arena layout ----
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 1023) & ~(size_t)1023; return o; };
  struct Tmp { size_t z, part, slab, small, dz, dzd, wT; };
  std::vector<Tmp> tmp(h->ops.size());
  T->layers.resize(h->ops.size());
  size_t flat = 0;
  for (size_t i = 0; i < h->ops.size(); ++i) {
    const Op& op = h->ops[i];
    TrainLayer& L = T->layers[i];
    L.Ho = (op.in.H + 2 * op.pad - op.kh) / op.stride + 1;
    L.Wo = (op.in.W + 2 * op.pad - op.kw) / op.stride + 1;
    L.has_bn = op.p_bn >= 0; L.has_bias = op.p_bias >= 0;
    const size_t K = (size_t)op.kh * op.kw * op.in.C;
    L.o_w = flat; flat += K * op.cout_pad;
    if (L.has_bn) { L.o_gamma = flat; flat += op.cout; L.o_beta = flat; flat += op.cout; }
    if (L.has_bias) { L.o_bias = flat; flat += op.cout; }
    flat = (flat + 3) & ~(size_t)3;
    const size_t Mmax = (size_t)B * L.Ho * L.Wo;
    L.stat_groups = umma_stats_groups((int)Mmax);
    tmp[i].small = take((size_t)(4 + 2 + 2 + 4) * op.cout * 4);                // mean, rstd, rmean, rvar | mg[2] | ab[2] | dzscale (padded)
    L.slab_cap = std::max(1, std::min(1024, kSlabCtas / ((op.cout + 255) / 256)));
    tmp[i].slab = take((size_t)L.slab_cap * 2 * op.cout * 8);
    if (L.has_bn) {
      if (op.out.buf < 0) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train_init: BatchNorm layer writing a user output"));
      L.z_ps = (long long)Mmax * op.cout;
      tmp[i].z = take((size_t)2 * L.z_ps * 2);
      tmp[i].part = take(L.stat_groups * 2 * op.cout * 4);
      const int g = (kGuardRows + L.Ho * L.Wo - 1) / (L.Ho * L.Wo);
      L.dz_plane_rows = (long long)(B + g) * L.Ho * L.Wo;
      tmp[i].dz = take((size_t)2 * L.dz_plane_rows * op.cout * 2);
      tmp[i].dzd = (size_t)-1;
      // 3x3 / stride 2 / pad 1 (every Darknet down-sampling conv): four parity-class convolutions over dz instead of one convolution
      // over a zero-dilated copy - a quarter of the MMAs and no dilated buffer (YOLO_B200_DGRAD_PARITY=0 keeps the dilated path)
      {
        const char* e = getenv("YOLO_B200_DGRAD_PARITY");
        L.dgrad_parity = op.stride == 2 && op.kh == 3 && op.kw == 3 && op.pad == 1 && op.in.buf >= 0 && op.in.H % 2 == 0 && op.in.W % 2 == 0 &&
                         op.cout % 32 == 0 && !umma_disabled() && !(e && e[0] == '0');
      }
      if (op.stride > 1 && op.in.buf >= 0 && !L.dgrad_parity) {
        const int gd = (kGuardRows + op.in.H * op.in.W - 1) / (op.in.H * op.in.W);
        L.dzd_n = B + gd;
        L.dzd_plane_rows = (long long)L.dzd_n * op.in.H * op.in.W;
        tmp[i].dzd = take((size_t)2 * L.dzd_plane_rows * op.cout * 2);
      }
      L.cin_pad = (op.in.C + 3) & ~3;
      // shapes the tensor-core data-gradient kernel does not take (Cout % 32 != 0) keep flipped fp32 weights for the FFMA kernel
      tmp[i].wT = (op.in.buf >= 0 && op.cout % 32 != 0) ? take((size_t)op.kh * op.kw * op.cout * L.cin_pad * 4) : (size_t)-1;
    } else {
      tmp[i].z = tmp[i].part = tmp[i].dz = tmp[i].dzd = (size_t)-1;
      L.cin_pad = (op.in.C + 3) & ~3;
      tmp[i].wT = op.in.buf >= 0 ? take((size_t)op.kh * op.kw * op.cout * L.cin_pad * 4) : (size_t)-1;
    }
  }
  if (flat != n_flat) return hfail(h, fail(YOLO_E_SHAPE, "train_init: internal flat-size mismatch %zu vs %zu", flat, n_flat));
  // head outputs + their gradients (the loss kernel writes dheads = dz of the head convs)
  size_t head_off[YOLO_MAX_SCALES + 1], dhead_off[YOLO_MAX_SCALES + 1];
  for (size_t i = 0; i < h->outputs.size(); ++i) {
    const View& v = h->outputs[i];
    head_off[i] = take((size_t)B * v.H * v.W * v.C * 4);
    dhead_off[i] = take((size_t)B * v.H * v.W * v.C * 4);
  }
  T->grads_begin = off;
  T->gbuf_off.resize(h->bufs.size());
  for (size_t i = 0; i < h->bufs.size(); ++i) {
    const size_t elems = h->bufs[i].bytes_per_image / dtype_bytes_per_elem(h->bufs[i].dtype);
    T->gbuf_off[i] = take(elems * (size_t)B * 4);
  }
  T->grads_bytes = off - T->grads_begin;
  const size_t lsb = yolo_loss_scratch_bytes(B, 16);
  const size_t o_ls = take(lsb);
  // weight-gradient scratch: the largest split-partial set of any layer (tensor-core kernel) / slab set (FFMA kernel)
  size_t wg_bytes = 0;
  for (size_t i = 0; i < h->ops.size(); ++i) {
    const Op& op = h->ops[i];
    const TrainLayer& L = T->layers[i];
    const size_t kn = (size_t)op.kh * op.kw * op.in.C * op.cout_pad * 4;
    const size_t Mmax = (size_t)B * L.Ho * L.Wo;
    size_t need = kn * (size_t)simt_wgrad_slabs((int)Mmax, op.kh * op.kw * op.in.C, op.cout_pad);     // FFMA slabs (the launcher's formula at max batch)
    if (L.has_bn && op.in.buf >= 0 && wgrad_umma_eligible(op.in.C, op.cout, op.kh, op.kw, op.in.dtype, op.in.il)) {
      const int n_units = op.kh * op.kw * (op.in.C / 64), bn = op.cout % 256 == 0 ? 256 : (op.cout % 128 == 0 ? 128 : 64);
      const int tiles = ((n_units + 1) / 2) * (op.cout / bn);
      int splits = tiles >= num_sms ? 1 : num_sms / tiles;
      const int nblk = (int)((Mmax + 63) / 64);
      splits = std::max(1, std::min(splits, nblk / 8));
      need = splits > 1 ? kn * splits : 0;
    }
    wg_bytes = std::max(wg_bytes, need);
  }
  const size_t o_wg = take(wg_bytes);
  if (cudaMalloc(reinterpret_cast<void**>(&T->arena), off) != cudaSuccess) {
    cudaGetLastError();
    delete T; h->train = nullptr;
    return hfail(h, fail(YOLO_E_OOM, "train_init: cudaMalloc(%zu) failed", off));
  }
  T->arena_bytes = off;
  T->loss_scratch = T->arena + o_ls; T->loss_scratch_bytes = lsb;
  T->wg_scratch = reinterpret_cast<float*>(T->arena + o_wg); T->wg_scratch_bytes = wg_bytes;
  YB_CUDA(cudaMemsetAsync(T->arena, 0, off, st));
  for (size_t i = 0; i < h->outputs.size(); ++i) T->dheads[i] = reinterpret_cast<float*>(T->arena + dhead_off[i]);
  // ---- fill the flat parameter buffer from the loaded parameters; running stats; per-layer plans ----
  std::vector<float> hostP(n_flat, 0.f);
  for (size_t i = 0; i < h->ops.size(); ++i) {
    Op& op = h->ops[i];
    TrainLayer& L = T->layers[i];
    const int cin = op.in.C, cout = op.cout;
    const float* Wsrc = h->params[op.p_weight].host.data();
    float wmax = 0.f;
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int r = 0; r < op.kh; ++r)
          for (int s2 = 0; s2 < op.kw; ++s2) {
            const float v = Wsrc[(((size_t)o * cin + c) * op.kh + r) * op.kw + s2];
            hostP[L.o_w + ((size_t)(r * op.kw + s2) * cin + c) * op.cout_pad + o] = v;
            wmax = fmaxf(wmax, fabsf(v));
          }
    if (L.has_bn) {
      memcpy(&hostP[L.o_gamma], h->params[op.p_bn].host.data(), cout * 4);
      memcpy(&hostP[L.o_beta], h->params[op.p_bn + 1].host.data(), cout * 4);
    }
    if (L.has_bias) memcpy(&hostP[L.o_bias], h->params[op.p_bias].host.data(), cout * 4);
    float* small = reinterpret_cast<float*>(T->arena + tmp[i].small);
    L.mean = small; L.rstd = small + cout; L.rmean = small + 2 * cout; L.rvar = small + 3 * cout;
    L.mg = small + 4 * cout; L.ab = small + 6 * cout; L.dzscale = small + 8 * cout;
    L.slab = reinterpret_cast<double*>(T->arena + tmp[i].slab);
    if (L.has_bn) {
      L.z = reinterpret_cast<__half*>(T->arena + tmp[i].z);
      L.stat_part = reinterpret_cast<float*>(T->arena + tmp[i].part);
      L.dz = reinterpret_cast<__half*>(T->arena + tmp[i].dz);
      L.dzd = tmp[i].dzd == (size_t)-1 ? nullptr : reinterpret_cast<__half*>(T->arena + tmp[i].dzd);
      YB_CUDA(cudaMemcpyAsync(L.rmean, h->params[op.p_bn + 2].host.data(), cout * 4, cudaMemcpyHostToDevice, st));
      YB_CUDA(cudaMemcpyAsync(L.rvar, h->params[op.p_bn + 3].host.data(), cout * 4, cudaMemcpyHostToDevice, st));
    } else {
      L.zf = reinterpret_cast<float*>(T->arena + head_off[-2 - op.out.buf]);
      L.wT = tmp[i].wT == (size_t)-1 ? nullptr : reinterpret_cast<float*>(T->arena + tmp[i].wT);
    }
    // weights may grow during training: pack with 2^11 of head-room below the fp16 range (inference packs at [256, 512))
    L.fwd_umma = op.umma.enabled;
    if (op.umma.eligible) umma_set_prescale(op.umma, wmax, 5);
    if (!op.w_f32_own) op.w_f32_own = op.w_f32;
    op.w_f32 = T->P + L.o_w;                                      // the FFMA kernels read the master weights directly
    if (L.has_bn && op.in.buf >= 0) {
      // data-gradient convolution over dz (or its zero-dilated copy): Cout -> Cin, stride 1, pad k-1-p, flipped filter
      const bool dil = op.stride > 1;
      if (L.dgrad_parity) {
        const int nimg = (int)(L.dz_plane_rows / ((long long)L.Ho * L.Wo));
        for (int q = 0; q < 4; ++q) {
          UmmaConv& u = L.dgrad_par[q];
          rc = umma_prepare_weights(u, YOLO_PREC_FP16X3, nullptr, cin, cout, 1 + (q >> 1), 1 + (q & 1), 1, 0, DT_F16X2, false, 0, false, st, true);
          if (rc) return hfail(h, rc);
          if (!u.eligible) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train_init: parity data gradient of layer %s is not a tensor-core shape", op.name.c_str()));
          umma_set_prescale(u, wmax, 5);
          u.pad_high_full = true;
          rc = umma_build_maps(u, L.dz, nimg, L.Ho, L.Wo, cout, cout, 0);
          if (rc) return hfail(h, rc);
          if (!u.enabled) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train_init: no tensor map for the parity data gradient of layer %s", op.name.c_str()));
        }
      } else {
      rc = umma_prepare_weights(L.dgrad, YOLO_PREC_FP16X3, nullptr, cin, cout, op.kh, op.kw, 1, op.kh - 1 - op.pad, DT_F16X2, false, 0, false, st);
      if (rc) return hfail(h, rc);
      }
      if (L.dgrad.eligible) {
        umma_set_prescale(L.dgrad, wmax, 5);
        const int Hd = dil ? op.in.H : L.Ho, Wd = dil ? op.in.W : L.Wo;
        const int nimg = dil ? L.dzd_n : (int)(L.dz_plane_rows / ((long long)L.Ho * L.Wo));
        rc = umma_build_maps(L.dgrad, dil ? (void*)L.dzd : (void*)L.dz, nimg, Hd, Wd, cout, cout, 0);
        if (rc) return hfail(h, rc);
      }
      if (!L.dgrad.enabled && !L.dgrad_parity) {                   // FFMA data gradient on flipped fp32 weights
        if (tmp[i].wT == (size_t)-1) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train_init: no data-gradient kernel for layer %s", op.name.c_str()));
        L.wT = reinterpret_cast<float*>(T->arena + tmp[i].wT);
      }
      if (wgrad_umma_eligible(cin, cout, op.kh, op.kw, op.in.dtype, op.in.il) ||
          wgrad_umma_il32_eligible(cin, cout, op.kh, op.kw, op.in.dtype, op.in.il, op.in.cpitch, op.in.coff)) {
        rc = wgrad_umma_plan(L.wg, act16(h, op.in), B, op.in.H, op.in.W, cin, op.in.cpitch, op.in.coff, op.kh, op.kw, op.stride, op.pad, L.dz,
                             L.dz_plane_rows, cout);
        if (rc) return hfail(h, rc);
      }
    }
  }
  // First-writer analysis of the fp32 gradient buffers: walking the ops in backward order, the first producer of a channel range
  // stores (no read-modify-write), later producers accumulate - and the 2.8 GB zero-fill per step disappears.  A range that is only
  // partly covered when somebody accumulates into it, or is read before anybody wrote it, falls back to zero-filling that buffer.
  {
    std::vector<std::vector<std::pair<int, int>>> written(h->bufs.size());
    T->gbuf_memset.assign(h->bufs.size(), 0);
    auto covered = [&](int buf, int lo, int hi) {                 // is [lo, hi) inside the union of the written ranges?
      std::vector<std::pair<int, int>> w = written[buf];
      std::sort(w.begin(), w.end());
      int pos = lo;
      for (auto& r : w) { if (r.first > pos) break; pos = std::max(pos, r.second); }
      return pos >= hi;
    };
    auto touches = [&](int buf, int lo, int hi) {
      for (auto& r : written[buf]) if (r.first < hi && lo < r.second) return true;
      return false;
    };
    auto write = [&](const View& v, bool& accum) {
      const int lo = v.coff, hi = v.coff + v.C;
      if (!touches(v.buf, lo, hi)) accum = false;
      else { accum = true; if (!covered(v.buf, lo, hi)) T->gbuf_memset[v.buf] = 1; }
      written[v.buf].push_back({lo, hi});
    };
    for (int i = (int)h->ops.size() - 1; i >= 0; --i) {
      const Op& op = h->ops[i];
      TrainLayer& L = T->layers[i];
      if (L.has_bn && op.out.buf >= 0 && !covered(op.out.buf, op.out.coff, op.out.coff + op.out.C)) T->gbuf_memset[op.out.buf] = 1;   // read before written
      if (L.has_bn && op.has_res) write(op.res, L.dres_accum);
      if (op.in.buf >= 0) write(op.in, L.dgrad_accum);
    }
    for (size_t i = 0; i < h->ops.size(); ++i) {                  // a zero-filled buffer is accumulated into by everybody
      const Op& op = h->ops[i];
      if (op.in.buf >= 0 && T->gbuf_memset[op.in.buf]) T->layers[i].dgrad_accum = true;
      if (op.has_res && T->gbuf_memset[op.res.buf]) T->layers[i].dres_accum = true;
    }
  }
  YB_CUDA(cudaMemcpyAsync(T->P, hostP.data(), n_flat * 4, cudaMemcpyHostToDevice, st));
  YB_CUDA(cudaMemsetAsync(T->M1, 0, n_flat * 4, st));
  YB_CUDA(cudaMemsetAsync(T->M2, 0, n_flat * 4, st));
  YB_CUDA(cudaMemsetAsync(T->G, 0, n_flat * 4, st));
  rc = repack_weights(h, st);
  if (rc) return hfail(h, rc);
  YB_CUDA(cudaStreamSynchronize(st));
  return YOLO_OK;
}

extern "C" int yolo_train_set_bn_momentum(yolo_handle* h, float momentum) {
  if (!h || !h->train) return fail(YOLO_E_STATE, "train_set_bn_momentum: yolo_train_init has not been called");
  if (!(momentum >= 0.f && momentum <= 1.f)) return hfail(h, fail(YOLO_E_BADARG, "train_set_bn_momentum: momentum must be in [0, 1]"));
  h->train->bn_momentum = momentum;
  return YOLO_OK;
}

// ---- data-parallel gradient exchange --------------------------------------------------------------------------------
extern "C" int yolo_nccl_unique_id(void* id128) {
  if (!id128) return fail(YOLO_E_BADARG, "nccl_unique_id: null argument");
  int rc = load_nccl();
  if (rc) return rc;
  YB_NCCL(g_nccl.GetUniqueId(id128));
  return YOLO_OK;
}

extern "C" int yolo_train_comm_init(yolo_handle* h, const void* id128, int rank, int world, size_t bucket_bytes) {
  if (!h || !h->train) return fail(YOLO_E_STATE, "train_comm_init: yolo_train_init has not been called");
  if (!id128 || world < 1 || rank < 0 || rank >= world) return hfail(h, fail(YOLO_E_BADARG, "train_comm_init: bad arguments"));
  TrainState* T = h->train;
  YB_CUDA(cudaSetDevice(h->device));
  int rc = load_nccl();
  if (rc) return hfail(h, rc);
  if (T->comm) { g_nccl.CommDestroy(T->comm); T->comm = nullptr; }
  NcclUid uid;
  memcpy(uid.b, id128, 128);
  int r = g_nccl.CommInitRank(&T->comm, world, uid, rank);
  if (r != 0) return hfail(h, fail(YOLO_E_NCCL, "ncclCommInitRank(rank %d of %d) -> %s", rank, world, g_nccl.GetErrorString(r)));
  T->world = world; T->rank = rank;
  if (bucket_bytes >= (1u << 20)) T->bucket_floats = bucket_bytes / 4;
  if (!T->comm_stream) YB_CUDA(cudaStreamCreateWithFlags(&T->comm_stream, cudaStreamNonBlocking));
  if (!T->ev_bucket) YB_CUDA(cudaEventCreateWithFlags(&T->ev_bucket, cudaEventDisableTiming));
  if (!T->ev_done) YB_CUDA(cudaEventCreateWithFlags(&T->ev_done, cudaEventDisableTiming));
  return YOLO_OK;
}

// forward (train-mode BN) + targets/losses + backward: fills the flat gradient buffer (d sum(losses) / d param); with a communicator
// the gradient is all-reduced (sum over ranks) bucket by bucket while the backward runs.
static int train_fwd_bwd(yolo_handle* h, const void* input, int in_layout, const float* labels, int batch, int n_obj, const yolo_loss_params* lp,
                         const float* lp_labels, int n_lp_obj, int n_lp_lab, const yolo_lp_loss_params* lpp, float* out_losses, void* stream) {
  if (!h || !h->train) return fail(YOLO_E_STATE, "train: yolo_train_init has not been called");
  if (!input || !labels || !lp || !out_losses) return hfail(h, fail(YOLO_E_BADARG, "train: null argument"));
  if (batch < 1 || batch > h->spec.max_batch) return hfail(h, fail(YOLO_E_SHAPE, "train: batch=%d outside [1,%d]", batch, h->spec.max_batch));
  if (n_obj > 16) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train: at most 16 labels per image"));
  YB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  TrainState* T = h->train;
  const int launches0 = g_launches;
  const int lay_in = in_layout == YOLO_IN_NCHW_F32 ? 1 : 2;
  const yolo_spec& s = h->spec;
  int rc;
  // ---------------- forward ----------------
  for (size_t i = 0; i < h->ops.size(); ++i) {
    const Op& op = h->ops[i];
    TrainLayer& L = T->layers[i];
    const int M = batch * L.Ho * L.Wo, C = op.cout;
    ConvDesc d;
    fill_conv_desc(h, op, batch, input, nullptr, d);
    d.scale = nullptr; d.act = ACT_NONE; d.res = nullptr; d.upsample2 = 0;
    d.w_f32 = T->P + L.o_w;
    if (L.has_bn) {
      d.shift = nullptr;
      d.out = L.z; d.out_dtype = DT_F16X2; d.out_cpitch = C; d.out_coff = 0; d.out_plane_stride = L.z_ps;
    } else {
      d.shift = L.has_bias ? T->P + L.o_bias : nullptr;          // head convs: conv + bias is the output itself
      d.out = L.zf; d.out_dtype = DT_F32; d.out_cpitch = C; d.out_coff = 0; d.out_plane_stride = 0;
    }
    const int lay = op.in.buf == -1 ? lay_in : 0;
    bool fused_stats = false;
    size_t groups = 0;
    if (op.umma.enabled) {
      UmmaExtra ex;
      if (L.has_bn) { ex.stats = L.stat_part; fused_stats = true; }
      rc = launch_conv_umma(op.umma, d, st, &ex);
      groups = ex.stats_groups_out;
    } else if (stem_eligible(d, lay)) rc = launch_stem(d, lay, st);
    else rc = launch_conv_simt(d, lay, st);
    if (rc) return hfail(h, rc);
    if (!L.has_bn) continue;
    if (!fused_stats) {
      groups = (size_t)(M + 31) / 32;
      const size_t warps = groups * ((C + 31) / 32);
      bn_stats_rows_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(L.z, L.z_ps, M, C, L.stat_part);
      ++g_launches;
    }
    const int nslab = (int)std::max<size_t>(1, std::min<size_t>(std::min(L.slab_cap, kSlabCtas / ((C + 31) / 32)), groups / 32));
    reduce_groups_kernel<<<dim3((C + 31) / 32, nslab), 256, 0, st>>>(L.stat_part, groups, C, L.slab);
    bn_finalize_fwd_kernel<<<(C + 7) / 8, 256, 0, st>>>(L.slab, nslab, M, C, L.mean, L.rstd, L.rmean, L.rvar, T->bn_momentum, T->P + L.o_gamma,
                                                            T->P + L.o_beta, L.ab);
    const __half* res = op.has_res ? act16(h, op.res) : nullptr;
    {
      const int cb = std::min(C, 256), lanes = 256 / (cb / 8);
      const int rowblocks = std::max(1, std::min((M + lanes - 1) / lanes, 148 * 8 / ((C + cb - 1) / cb)));
      bn_act_fwd_kernel<<<dim3((C + cb - 1) / cb, rowblocks), 256, 0, st>>>(L.z, L.z_ps, M, C, L.ab, op.act, res, op.res.cpitch, op.res.coff, op.res.ps,
                                                                          act16(h, op.out), op.out.cpitch, op.out.coff, op.out.ps, op.upsample2, L.Ho,
                                                                          L.Wo, h->d_flags);
    }
    g_launches += 3;
  }
  YB_CUDA(cudaGetLastError());
  // ---------------- targets, losses, d loss / d heads ----------------
  yolo_decode_geom g;
  memset(&g, 0, sizeof(g));
  g.height = s.height; g.width = s.width; g.n_scales = s.n_scales; g.n_anchors = s.n_anchors; g.channels_per_anchor = s.channels_per_anchor;
  for (int i = 0; i < s.n_scales; ++i) {
    g.step[i] = 1 << (s.n_layers - s.n_scales + 1 + i);
    for (int a = 0; a < s.n_anchors; ++a) { g.anchors[i][a][0] = s.anchors[i][a][0]; g.anchors[i][a][1] = s.anchors[i][a][1]; }
  }
  const void* heads[YOLO_MAX_SCALES + 1];
  void* dheads[YOLO_MAX_SCALES + 1];
  for (size_t i = 0; i < h->ops.size(); ++i)
    if (h->ops[i].out.buf < -1) heads[-2 - h->ops[i].out.buf] = T->layers[i].zf;
  for (int i = 0; i < s.n_scales; ++i) dheads[i] = T->dheads[i];
  rc = yolo_loss_targets(&g, heads, labels, batch, n_obj, lp, T->loss_scratch, out_losses, dheads, nullptr, stream);
  if (rc) return hfail(h, rc);
  for (size_t i = s.n_scales; i < h->outputs.size(); ++i) {
    const View& v = h->outputs[i];
    if (lp_labels) {                                              // car_and_LP `_train_batch`: the five LP losses join the backward
      rc = yolo_lp_loss_targets(static_cast<const float*>(heads[i]), 0, batch, v.H, v.W, v.C, s.height / v.H, s.lp_r_max, lp_labels, n_lp_obj, n_lp_lab,
                                lpp, out_losses + (size_t)5 * batch, T->dheads[i], stream);
      if (rc) return hfail(h, rc);
    } else                                                        // car/YOLO.py `_train_batch` on a CarLPNet: no loss on the LP map, zero gradient
      YB_CUDA(cudaMemsetAsync(T->dheads[i], 0, (size_t)batch * v.H * v.W * v.C * 4, st));
  }
  // ---------------- backward ----------------
  for (size_t bi = 0; bi < h->bufs.size(); ++bi)
    if (T->gbuf_memset[bi]) {
      const size_t elems = h->bufs[bi].bytes_per_image / dtype_bytes_per_elem(h->bufs[bi].dtype);
      YB_CUDA(cudaMemsetAsync(T->arena + T->gbuf_off[bi], 0, elems * (size_t)h->spec.max_batch * 4, st));
    }
  T->reduced = false;
  size_t bucket_hi = T->n_flat;                                   // gradients of [bucket_lo, bucket_hi) are complete when the walk passes bucket_lo
  for (int i = (int)h->ops.size() - 1; i >= 0; --i) {
    const Op& op = h->ops[i];
    TrainLayer& L = T->layers[i];
    const int M = batch * L.Ho * L.Wo, C = op.cout;
    const int K = op.kh * op.kw * op.in.C;
    const int lay = op.in.buf == -1 ? lay_in : 0;
    if (L.has_bn) {
      const float* dy = grad_ptr(h, op.out);
      const int dyp = grad_pitch(op.out);
      float* dres = op.has_res ? grad_ptr(h, op.res) : nullptr;
      YB_CUDA(cudaMemsetAsync(L.dzscale + 2, 0, 8, st));
      const int cb = std::min(C, 256);
      const int red_lanes = 256 / (cb / 8);
      const int nslab = std::max(1, std::min(std::min(L.slab_cap, kSlabCtas / ((C + cb - 1) / cb)), M / (red_lanes * 16)));
      bn_bwd_reduce_kernel<<<dim3((C + cb - 1) / cb, nslab), 256, 0, st>>>(L.z, L.z_ps, M, C, dy, dyp, op.out.coff, op.upsample2, L.Ho, L.Wo,
                                                                         L.ab, op.act, dres, op.has_res ? grad_pitch(op.res) : 0, op.res.coff,
                                                                         L.dres_accum ? 1 : 0, L.slab, reinterpret_cast<unsigned int*>(L.dzscale + 2));
      bn_bwd_finalize_kernel<<<(C + 7) / 8, 256, 0, st>>>(L.slab, nslab, M, C, L.mean, L.rstd, L.ab, L.mg, T->G + L.o_gamma, T->G + L.o_beta,
                                                          reinterpret_cast<unsigned int*>(L.dzscale + 3));
      {
        const int lanes = 256 / (cb / 8);
        const int rowblocks = std::max(1, std::min((M + kGuardRows + lanes - 1) / lanes, 148 * 8 / ((C + cb - 1) / cb)));
        bn_bwd_apply_kernel<<<dim3((C + cb - 1) / cb, rowblocks), 256, 0, st>>>(
            L.z, L.z_ps, M, C, L.mg, dy, dyp, op.out.coff, op.upsample2, L.Ho, L.Wo, L.mean, L.rstd, L.ab, op.act, L.dzscale, L.dz, L.dz_plane_rows * C,
            L.dzd, L.dzd_plane_rows * C, op.stride, op.in.H, op.in.W, h->d_flags);
      }
      g_launches += 3;
      // weight gradient
      if (L.wg.enabled) {
        rc = launch_wgrad_umma(L.wg, batch, T->G + L.o_w, op.cout_pad, 1.f, L.dzscale + 1, T->wg_scratch, T->wg_scratch_bytes, 0, st);
        if (rc) return hfail(h, rc);
      } else {
        int slabs = simt_wgrad_slabs(M, K, op.cout_pad);
        const size_t kn = (size_t)K * op.cout_pad;
        dim3 gw((K + 63) / 64, (op.cout_pad + 63) / 64, slabs);
        const void* xin = op.in.buf == -1 ? input : (const void*)act16(h, op.in);
        if (op.in.buf == -1 && op.in.C == 3 && K <= 32 && C <= 32 && op.cout_pad == C && (size_t)K * C <= 1024) {
          slabs = std::max(1, std::min(M / 256, kSlabCtas));
          if ((size_t)slabs * kn * 4 > T->wg_scratch_bytes) slabs = std::max(1, (int)(T->wg_scratch_bytes / (kn * 4)));
          stem_wgrad_kernel<<<slabs, 256, 0, st>>>(xin, lay, batch, op.in.H, op.in.W, L.dz, L.dz_plane_rows * C, L.Ho, L.Wo, C, op.kh, op.kw, op.stride,
                                                   op.pad, T->wg_scratch, kn, op.cout_pad);
        } else
        wgrad_simt_kernel<true, true><<<gw, 256, 0, st>>>(xin, op.in.ps, batch, op.in.H, op.in.W, op.in.C, op.in.cpitch, op.in.coff, lay, L.dz,
                                                         L.dz_plane_rows * C, C, L.Ho, L.Wo, C, op.kh, op.kw, op.stride, op.pad, T->wg_scratch, kn,
                                                         op.cout_pad);
        wgrad_reduce_scaled_kernel<<<grid_for(kn), 256, 0, st>>>(T->wg_scratch, slabs, kn, kn, L.dzscale + 1, T->G + L.o_w);
        g_launches += 2;
      }
      // data gradient (not needed for the network input): accumulate into the fp32 gradient buffer of the input view
      if (op.in.buf >= 0) {
        ConvDesc d;
        memset(&d, 0, sizeof(d));
        const bool dil = op.stride > 1;
        d.in = dil ? (const void*)L.dzd : (const void*)L.dz; d.in_dtype = DT_F16X2; d.N = batch;
        d.H = dil ? op.in.H : L.Ho; d.W = dil ? op.in.W : L.Wo; d.Cin = C; d.in_cpitch = C; d.in_coff = 0;
        d.in_plane_stride = (dil ? L.dzd_plane_rows : L.dz_plane_rows) * C;
        d.kh = op.kh; d.kw = op.kw; d.stride = 1; d.pad = op.kh - 1 - op.pad; d.Cout = op.in.C;
        d.act = ACT_NONE;
        d.out = grad_ptr(h, op.in); d.out_dtype = DT_F32; d.Ho = op.in.H; d.Wo = op.in.W; d.out_cpitch = grad_pitch(op.in); d.out_coff = op.in.coff;
        d.sat_flag = h->d_flags;
        if (L.dgrad_parity) {
          d.in = L.dz; d.H = L.Ho; d.W = L.Wo; d.in_plane_stride = L.dz_plane_rows * C; d.pad = 0; d.Ho = L.Ho; d.Wo = L.Wo;
          for (int q = 0; q < 4 && !rc; ++q) {
            d.kh = 1 + (q >> 1); d.kw = 1 + (q & 1);
            d.upsample2 = 2 + q;                                      // epilogue: pixel (2a + py, 2b + px) of the input gradient
            UmmaExtra ex;
            ex.acc_scale_dev = L.dzscale + 1; ex.accum = L.dgrad_accum ? 1 : 0;
            rc = launch_conv_umma(L.dgrad_par[q], d, st, &ex);
          }
        } else if (L.dgrad.enabled) {
          UmmaExtra ex;
          ex.acc_scale_dev = L.dzscale + 1; ex.accum = L.dgrad_accum ? 1 : 0;
          rc = launch_conv_umma(L.dgrad, d, st, &ex);
        } else {
          d.in = L.dz; d.H = L.Ho; d.W = L.Wo; d.in_plane_stride = L.dz_plane_rows * C; d.in_dil = op.stride;      // implicit zero-dilation
          d.w_f32 = L.wT; d.cout_pad = L.cin_pad;
          d.dyn_scale = L.dzscale + 1;
          if (L.dgrad_accum) { d.res = d.out; d.res_cpitch = d.out_cpitch; d.res_coff = d.out_coff; }
          rc = launch_conv_simt(d, 0, st);
        }
        if (rc) return hfail(h, rc);
      }
    } else {
      // head conv (no BatchNorm, 90 / 10 channels): fp32 FFMA kernels on dz = d loss / d head
      float* dz = T->dheads[-2 - op.out.buf];
      if (L.has_bias) {
        const int nslab = std::max(1, std::min(std::min(L.slab_cap, kSlabCtas / ((C + 31) / 32)), M / 64));
        colsum_slab_kernel<<<dim3((C + 31) / 32, nslab), 256, 0, st>>>(dz, M, C, L.slab);
        colsum_finalize_kernel<<<(C + 7) / 8, 256, 0, st>>>(L.slab, nslab, C, T->G + L.o_bias);
        ++g_launches;
      }
      const int slabs = simt_wgrad_slabs(M, K, op.cout_pad);
      const size_t kn = (size_t)K * op.cout_pad;
      dim3 gw((K + 63) / 64, (op.cout_pad + 63) / 64, slabs);
      wgrad_simt_kernel<true, false><<<gw, 256, 0, st>>>(act16(h, op.in), op.in.ps, batch, op.in.H, op.in.W, op.in.C, op.in.cpitch, op.in.coff, 0, dz, 0, C,
                                                        L.Ho, L.Wo, C, op.kh, op.kw, op.stride, op.pad, T->wg_scratch, kn, op.cout_pad);
      wgrad_reduce_scaled_kernel<<<grid_for(kn), 256, 0, st>>>(T->wg_scratch, slabs, kn, kn, nullptr, T->G + L.o_w);
      g_launches += 3;
      ConvDesc d;
      memset(&d, 0, sizeof(d));
      d.in = dz; d.in_dtype = DT_F32; d.N = batch; d.H = L.Ho; d.W = L.Wo; d.Cin = C; d.in_cpitch = C; d.in_coff = 0;
      d.kh = op.kh; d.kw = op.kw; d.stride = 1; d.pad = op.kh - 1 - op.pad; d.in_dil = op.stride; d.Cout = op.in.C;
      d.w_f32 = L.wT; d.cout_pad = L.cin_pad;
      d.act = ACT_NONE;
      float* dx = grad_ptr(h, op.in);
      if (L.dgrad_accum) { d.res = dx; d.res_cpitch = grad_pitch(op.in); d.res_coff = op.in.coff; }      // several consumers may feed one tensor
      d.out = dx; d.out_dtype = DT_F32; d.Ho = op.in.H; d.Wo = op.in.W; d.out_cpitch = grad_pitch(op.in); d.out_coff = op.in.coff;
      rc = launch_conv_simt(d, 0, st);
      if (rc) return hfail(h, rc);
    }
    // gradient exchange: release the bucket that ends at this layer once it is large enough (or the walk is done)
    if (T->comm && T->world > 1) {
      const size_t lo = L.o_w;
      if (bucket_hi - lo >= T->bucket_floats || i == 0) {
        YB_CUDA(cudaEventRecord(T->ev_bucket, st));
        YB_CUDA(cudaStreamWaitEvent(T->comm_stream, T->ev_bucket, 0));
        int r = g_nccl.AllReduce(T->G + lo, T->G + lo, bucket_hi - lo, /*ncclFloat32*/ 7, /*ncclSum*/ 0, T->comm, T->comm_stream);
        if (r != 0) return hfail(h, fail(YOLO_E_NCCL, "ncclAllReduce(bucket [%zu, %zu)) -> %s", lo, bucket_hi, g_nccl.GetErrorString(r)));
        bucket_hi = lo;
      }
    }
  }
  if (T->comm && T->world > 1) {
    YB_CUDA(cudaEventRecord(T->ev_done, T->comm_stream));
    YB_CUDA(cudaStreamWaitEvent(st, T->ev_done, 0));             // the update (and any reader of G on `st`) sees the summed gradient
    T->reduced = true;
  }
  YB_CUDA(cudaGetLastError());
  h->last_launches = g_launches - launches0;
  return YOLO_OK;
}

extern "C" int yolo_train_forward_backward(yolo_handle* h, const void* input, int in_layout, const float* labels, int batch, int n_obj,
                                           const yolo_loss_params* lp, float* out_losses, void* stream) {
  return train_fwd_bwd(h, input, in_layout, labels, batch, n_obj, lp, nullptr, 0, 0, nullptr, out_losses, stream);
}

extern "C" int yolo_train_forward_backward_lp(yolo_handle* h, const void* input, int in_layout, const float* labels, int batch, int n_obj,
                                              const yolo_loss_params* lp, const float* lp_labels, int n_lp_obj, int n_lp_lab,
                                              const yolo_lp_loss_params* lpp, float* out_losses, void* stream) {
  if (!h) return fail(YOLO_E_BADARG, "train: null handle");
  if (h->spec.net_type != YOLO_NET_CARLPNET) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train: LP losses need a CARLPNET"));
  if (!lp_labels || !lpp) return hfail(h, fail(YOLO_E_BADARG, "train: null LP labels / parameters"));
  return train_fwd_bwd(h, input, in_layout, labels, batch, n_obj, lp, lp_labels, n_lp_obj, n_lp_lab, lpp, out_losses, stream);
}

// trainer.step(batch_size): G holds the SUM over ranks (reduced during the backward when a communicator is attached, else by the
// caller); rescale = 1/batch_size
extern "C" int yolo_train_apply(yolo_handle* h, float lr, float beta1, float beta2, float eps, float rescale_grad, void* stream) {
  if (!h || !h->train) return fail(YOLO_E_STATE, "train_apply: yolo_train_init has not been called");
  YB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  TrainState* T = h->train;
  const int launches0 = g_launches;
  const int t = ++T->step_count;
  const float lr_t = lr * sqrtf(1.f - powf(beta2, (float)t)) / (1.f - powf(beta1, (float)t));     // mxnet.optimizer.Adam
  adam_kernel<<<grid_for(T->n_flat / 4, 256, 148 * 16), 256, 0, st>>>(reinterpret_cast<float4*>(T->P), reinterpret_cast<const float4*>(T->G),
                                                                      reinterpret_cast<float4*>(T->M1), reinterpret_cast<float4*>(T->M2), T->n_flat / 4, lr_t,
                                                                      rescale_grad, beta1, beta2, eps);
  ++g_launches;
  int rc = repack_weights(h, st);                                 // fp16 planes of both convolution directions follow the master weights
  if (rc) return hfail(h, rc);
  for (size_t i = 0; i < h->ops.size(); ++i) {                    // keep the inference epilogues consistent with the trained parameters
    Op& op = h->ops[i];
    TrainLayer& L = T->layers[i];
    if (!op.scale) continue;
    refold_kernel<<<(op.cout + 127) / 128, 128, 0, st>>>(L.has_bn ? T->P + L.o_gamma : nullptr, L.has_bn ? T->P + L.o_beta : nullptr, L.rmean, L.rvar,
                                                         L.has_bias ? T->P + L.o_bias : nullptr, op.cout, op.scale, op.shift);
    ++g_launches;
  }
  YB_CUDA(cudaGetLastError());
  h->last_launches += g_launches - launches0;
  return YOLO_OK;
}

// read a parameter (or running statistic) back in the canonical layout of yolo_load_param
extern "C" int yolo_get_param(yolo_handle* h, const char* name, float* host, size_t n_elems, int want_grad) {
  if (!h || !name || !host) return fail(YOLO_E_BADARG, "get_param: null argument");
  auto it = h->pindex.find(name);
  if (it == h->pindex.end()) return hfail(h, fail(YOLO_E_BADARG, "get_param: unknown parameter '%s'", name));
  const Param& p = h->params[it->second];
  if (p.numel() != n_elems) return hfail(h, fail(YOLO_E_SHAPE, "get_param: '%s' has %zu elements", name, p.numel()));
  if (!h->train) {
    if (want_grad) return hfail(h, fail(YOLO_E_STATE, "get_param: no gradients before yolo_train_init"));
    memcpy(host, p.host.data(), n_elems * 4);
    return YOLO_OK;
  }
  YB_CUDA(cudaSetDevice(h->device));
  YB_CUDA(cudaDeviceSynchronize());
  TrainState* T = h->train;
  const float* base = want_grad ? T->G : T->P;
  for (size_t i = 0; i < h->ops.size(); ++i) {
    const Op& op = h->ops[i];
    const TrainLayer& L = T->layers[i];
    const int idx = it->second;
    if (idx == op.p_weight) {
      const int cin = op.in.C, cout = op.cout;
      std::vector<float> tmp((size_t)op.kh * op.kw * cin * op.cout_pad);
      YB_CUDA(cudaMemcpy(tmp.data(), base + L.o_w, tmp.size() * 4, cudaMemcpyDeviceToHost));
      for (int o = 0; o < cout; ++o)
        for (int c = 0; c < cin; ++c)
          for (int r = 0; r < op.kh; ++r)
            for (int s2 = 0; s2 < op.kw; ++s2)
              host[(((size_t)o * cin + c) * op.kh + r) * op.kw + s2] = tmp[((size_t)(r * op.kw + s2) * cin + c) * op.cout_pad + o];
      return YOLO_OK;
    }
    if (op.p_bias >= 0 && idx == op.p_bias) { YB_CUDA(cudaMemcpy(host, base + L.o_bias, n_elems * 4, cudaMemcpyDeviceToHost)); return YOLO_OK; }
    if (op.p_bn >= 0 && idx >= op.p_bn && idx < op.p_bn + 4) {
      const int q = idx - op.p_bn;
      if (q >= 2 && want_grad) return hfail(h, fail(YOLO_E_BADARG, "get_param: running statistics have no gradient"));
      const float* src = q == 0 ? base + L.o_gamma : (q == 1 ? base + L.o_beta : (q == 2 ? L.rmean : L.rvar));
      YB_CUDA(cudaMemcpy(host, src, n_elems * 4, cudaMemcpyDeviceToHost));
      return YOLO_OK;
    }
  }
  return hfail(h, fail(YOLO_E_BADARG, "get_param: '%s' is not attached to an op", name));
}
