// One data-parallel training step of the YOLO net on one GPU, fp32.
//
// Replaces (reference, file:line): `_train_batch` car/YOLO.py:350-399 - forward with batch-statistics BatchNorm
// (per device, num_sync_bn_devices=-1 :94-96), `_loss_mask`/`_get_loss` (-> train_loss.cu), `sum(losses).backward()` :394
// (cuDNN bwd-data / bwd-filter / BN backward dispatched by MXNet autograd) and `trainer.step(batch_size)` :396 (kvstore
// reduce + `adam_update`).  Here: a forward that keeps the pre-BN conv outputs, a hand-written backward over the same flat
// op list (BN backward with double-precision channel reductions, LeakyReLU', residual / upsample / concat gradient routing
// through the same channel-slice views as the forward, data-gradient as a convolution with flipped weights on the FFMA
// kernel, weight-gradient as a split-M outer-product kernel), and a fused rescale + Adam update on ONE flat parameter /
// gradient buffer (caller-owned, so the gradient all-reduce between ranks is a single torch.distributed / NCCL call).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "net_internal.cuh"

namespace yb {

constexpr float kBnMomentum = 0.9f;       // gluon BatchNorm() default
constexpr float kTrainBnEps = 1e-5f;

struct TrainLayer {                       // per conv op
  // offsets (in floats) into the flat parameter / gradient buffers
  size_t o_w = 0, o_gamma = 0, o_beta = 0, o_bias = 0;
  bool has_bn = false, has_bias = false;
  float* z = nullptr;                     // pre-BN conv output, dense [max_batch*Ho*Wo][Cout] (reused as dz in the backward)
  float* rmean = nullptr; float* rvar = nullptr;       // running statistics (device)
  double* sums = nullptr;                 // [4][Cout]: sum, sumsq (forward) / sum g, sum g*xhat (backward)
  float* mean = nullptr; float* rstd = nullptr;
  float* wT = nullptr;                    // flipped/transposed weights for the data gradient [kh*kw*Cout][cin_pad]
  int cin_pad = 0;
  int Ho = 0, Wo = 0;
};

struct TrainState {
  float* P = nullptr; float* G = nullptr; float* M1 = nullptr; float* M2 = nullptr;    // caller-owned flat buffers
  size_t n_flat = 0;
  std::vector<TrainLayer> layers;
  char* arena = nullptr;                  // cudaMalloc: z buffers, stats, wT, activation-gradient buffers
  size_t arena_bytes = 0;
  std::vector<size_t> gbuf_off;           // byte offset of the gradient buffer mirroring forward buffer i
  size_t grads_begin = 0, grads_bytes = 0;
  void* loss_scratch = nullptr; size_t loss_scratch_bytes = 0;
  float* dheads[YOLO_MAX_SCALES + 1] = {nullptr, nullptr, nullptr, nullptr};
  int step_count = 0;
};

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
// per-channel sums over the rows of a dense [M][C] matrix (two quantities at once), double accumulation
template <int MODE>   // 0: (z, z*z)   1: backward (g, g*xhat) with g = dy * leaky'(u); also routes dy into the residual's gradient
__global__ void __launch_bounds__(256)
channel_sums_kernel(const float* __restrict__ z, int M, int C, double* __restrict__ sums,
                    // backward only:
                    const float* __restrict__ dy, int dy_cpitch, int dy_coff, int upsample2, int Ho, int Wo,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, int act, float* __restrict__ dres, int dres_cpitch, int dres_coff) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;                       // 8 row lanes
  const int rows_per_block = (M + gridDim.y - 1) / gridDim.y;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  double s0 = 0.0, s1 = 0.0;
  if (c < C) {
    float mu = 0.f, rs = 0.f, ga = 0.f, be = 0.f;
    if (MODE == 1) { mu = mean[c]; rs = rstd[c]; ga = gamma[c]; be = beta[c]; }
    for (int m = m0 + rl; m < m1; m += 8) {
      const float zv = z[(size_t)m * C + c];
      if (MODE == 0) { s0 += zv; s1 += (double)zv * zv; }
      else {
        float g;
        if (upsample2) {
          const int HoWo = Ho * Wo, n = m / HoWo, rem = m - n * HoWo, oh = rem / Wo, ow = rem - oh * Wo;
          g = 0.f;
          for (int q = 0; q < 4; ++q) {
            size_t pix = ((size_t)n * 2 * Ho + 2 * oh + (q >> 1)) * (2 * Wo) + 2 * ow + (q & 1);
            g += dy[pix * dy_cpitch + dy_coff + c];
          }
        } else g = dy[(size_t)m * dy_cpitch + dy_coff + c];
        if (dres) dres[(size_t)m * dres_cpitch + dres_coff + c] += g;       // y = act(bn(z)) + res  ->  d res += dy
        const float xh = (zv - mu) * rs;
        const float u = fmaf(xh, ga, be);
        if (act == ACT_LEAKY) g = u > 0.f ? g : 0.1f * g;
        else if (act == ACT_RELU) g = u > 0.f ? g : 0.f;
        s0 += g; s1 += (double)g * xh;
      }
    }
  }
  __shared__ double sh[2][8][32];
  sh[0][rl][threadIdx.x & 31] = s0; sh[1][rl][threadIdx.x & 31] = s1;
  __syncthreads();
  if (rl == 0 && c < C) {
    for (int r = 1; r < 8; ++r) { s0 += sh[0][r][threadIdx.x & 31]; s1 += sh[1][r][threadIdx.x & 31]; }
    atomicAdd(&sums[c], s0);
    atomicAdd(&sums[C + c], s1);
  }
}

__global__ void bn_finalize_fwd_kernel(const double* __restrict__ sums, int M, int C, float* mean, float* rstd, float* rmean, float* rvar) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = sums[c] / M, var = fmax(sums[C + c] / M - mu * mu, 0.0);
  mean[c] = (float)mu;
  rstd[c] = (float)(1.0 / sqrt(var + (double)kTrainBnEps));
  rmean[c] = rmean[c] * kBnMomentum + (float)mu * (1.f - kBnMomentum);
  rvar[c] = rvar[c] * kBnMomentum + (float)var * (1.f - kBnMomentum);
}

// y = act(gamma * (z - mean) * rstd + beta) (+ residual), written through the op's output view (concat slice / 2x upsample)
__global__ void bn_act_fwd_kernel(const float* __restrict__ z, int M, int C, const float* __restrict__ mean, const float* __restrict__ rstd,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, int act, const float* __restrict__ res,
                                  int res_cpitch, int res_coff, float* __restrict__ out, int out_cpitch, int out_coff, int upsample2, int Ho,
                                  int Wo) {
  const size_t total = (size_t)M * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int m = (int)(i / C);
    float u = fmaf((z[i] - mean[c]) * rstd[c], gamma[c], beta[c]);
    if (act == ACT_LEAKY) u = u > 0.f ? u : 0.1f * u;
    else if (act == ACT_RELU) u = fmaxf(u, 0.f);
    if (res) u += res[(size_t)m * res_cpitch + res_coff + c];
    if (upsample2) {
      const int HoWo = Ho * Wo, n = m / HoWo, rem = m - n * HoWo, oh = rem / Wo, ow = rem - oh * Wo;
      for (int q = 0; q < 4; ++q) {
        size_t pix = ((size_t)n * 2 * Ho + 2 * oh + (q >> 1)) * (2 * Wo) + 2 * ow + (q & 1);
        out[pix * out_cpitch + out_coff + c] = u;
      }
    } else out[(size_t)m * out_cpitch + out_coff + c] = u;
  }
}

// dz = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)), in place over z; g recomputed from dy like in the reduction pass
__global__ void bn_bwd_apply_kernel(float* __restrict__ z, int M, int C, const double* __restrict__ sums, const float* __restrict__ dy,
                                    int dy_cpitch, int dy_coff, int upsample2, int Ho, int Wo, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta, int act) {
  const size_t total = (size_t)M * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int m = (int)(i / C);
    float g;
    if (upsample2) {
      const int HoWo = Ho * Wo, n = m / HoWo, rem = m - n * HoWo, oh = rem / Wo, ow = rem - oh * Wo;
      g = 0.f;
      for (int q = 0; q < 4; ++q) {
        size_t pix = ((size_t)n * 2 * Ho + 2 * oh + (q >> 1)) * (2 * Wo) + 2 * ow + (q & 1);
        g += dy[pix * dy_cpitch + dy_coff + c];
      }
    } else g = dy[(size_t)m * dy_cpitch + dy_coff + c];
    const float xh = (z[i] - mean[c]) * rstd[c];
    const float u = fmaf(xh, gamma[c], beta[c]);
    if (act == ACT_LEAKY) g = u > 0.f ? g : 0.1f * g;
    else if (act == ACT_RELU) g = u > 0.f ? g : 0.f;
    const float mg = (float)(sums[c] / M), mgx = (float)(sums[C + c] / M);
    z[i] = gamma[c] * rstd[c] * (g - mg - xh * mgx);
  }
}

__global__ void bn_param_grads_kernel(const double* __restrict__ sums, int C, float* dgamma, float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] = (float)sums[c];
  dgamma[c] = (float)sums[C + c];
}
__global__ void bias_grad_kernel(const double* __restrict__ sums, int C, float* dbias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) dbias[c] = (float)sums[c];
}

// dW[k][n] += sum_m A[m][k] * dz[m][n], A = im2col gather of the layer input; 64x64 output tile per CTA, split over M
__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ x, int N, int H, int W, int Cin, int x_cpitch, int x_coff, int in_layout, const float* __restrict__ dz,
             int Ho, int Wo, int Cout, int kh, int kw, int stride, int pad, float* __restrict__ dW, int cout_pad) {
  constexpr int TK = 64, TN = 64, TM = 16;
  __shared__ float As[TM][TK + 1];
  __shared__ float Ds[TM][TN];
  const int K = kh * kw * Cin, M = N * Ho * Wo, HoWo = Ho * Wo;
  const int k0 = blockIdx.x * TK, n0 = blockIdx.y * TN;
  const int slab = (M + gridDim.z - 1) / gridDim.z;
  const int m_begin = blockIdx.z * slab, m_end = min(M, m_begin + slab);
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
          if (in_layout == 1) v = x[((size_t)(n * Cin + c) * H + ih) * W + iw];                         // network input NCHW fp32
          else if (in_layout == 2) v = (float)reinterpret_cast<const unsigned char*>(x)[((size_t)(n * H + ih) * W + iw) * Cin + c] / 255.f;
          else v = x[((size_t)(n * H + ih) * W + iw) * x_cpitch + x_coff + c];
        }
      }
      As[ml][kl] = v;
    }
    for (int e = tid; e < TM * TN; e += 256) {
      const int ml = e / TN, nl = e - ml * TN;
      const int m = mb + ml, n = n0 + nl;
      Ds[ml][nl] = (m < m_end && n < Cout) ? dz[(size_t)m * Cout + n] : 0.f;
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + tk * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn * 4 + j;
      if (n < Cout) atomicAdd(&dW[(size_t)k * cout_pad + n], acc[i][j]);
    }
  }
}

// Same contraction with a 128 x 128 output tile, 8 x 8 outputs per thread, float4 gathers (NHWC fp32 input, Cin % 4 == 0,
// Cout % 4 == 0) and register prefetch of the next pixel chunk: 16 FMAs per shared-memory load instead of 2.
__global__ void __launch_bounds__(256)
wgrad128_kernel(const float* __restrict__ x, int N, int H, int W, int Cin, int x_cpitch, int x_coff, const float* __restrict__ dz,
                int Ho, int Wo, int Cout, int kh, int kw, int stride, int pad, float* __restrict__ dW, int cout_pad) {
  constexpr int TK = 128, TN = 128, TM = 16;
  __shared__ __align__(16) float As[TM][TK];
  __shared__ __align__(16) float Ds[TM][TN];
  const int K = kh * kw * Cin, M = N * Ho * Wo, HoWo = Ho * Wo;
  const int k0 = blockIdx.x * TK, n0 = blockIdx.y * TN;
  const int slab = (M + gridDim.z - 1) / gridDim.z;
  const int m_begin = blockIdx.z * slab, m_end = min(M, m_begin + slab);
  const int tid = threadIdx.x, tk = tid >> 4, tn = tid & 15;        // 16 x 16 threads, 8 x 8 outputs each (two 4-wide halves)
  // loader role: float4 column lc of pixel rows lr and lr + 8 of every chunk; the k / n decode is loop invariant
  const int lr = tid >> 5, lc = (tid & 31) * 4;
  const int k = k0 + lc, nn = n0 + lc;
  const bool k_ok = k < K, n_ok = nn < Cout;
  int r = 0, sx = 0, c = 0;
  if (k_ok) { const int tap = k / Cin; c = k - tap * Cin; r = tap / kw; sx = tap - r * kw; }
  auto gather = [&](int m, float4& a, float4& d) {
    a = make_float4(0.f, 0.f, 0.f, 0.f);
    d = a;
    if (m < m_end) {
      if (k_ok) {
        const int n = m / HoWo, rem = m - n * HoWo, oh = rem / Wo, ow = rem - oh * Wo;
        const int ih = oh * stride - pad + r, iw = ow * stride - pad + sx;
        if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W)
          a = __ldg(reinterpret_cast<const float4*>(x + ((size_t)(n * H + ih) * W + iw) * x_cpitch + x_coff + c));
      }
      if (n_ok) d = __ldg(reinterpret_cast<const float4*>(dz + (size_t)m * Cout + nn));
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float4 pa[2], pd[2];
  gather(m_begin + lr, pa[0], pd[0]);
  gather(m_begin + lr + 8, pa[1], pd[1]);
  for (int mb = m_begin; mb < m_end; mb += TM) {
    *reinterpret_cast<float4*>(&As[lr][lc]) = pa[0];
    *reinterpret_cast<float4*>(&As[lr + 8][lc]) = pa[1];
    *reinterpret_cast<float4*>(&Ds[lr][lc]) = pd[0];
    *reinterpret_cast<float4*>(&Ds[lr + 8][lc]) = pd[1];
    __syncthreads();
    if (mb + TM < m_end) {                                          // prefetch the next chunk while this one is multiplied
      gather(mb + TM + lr, pa[0], pd[0]);
      gather(mb + TM + lr + 8, pa[1], pd[1]);
    }
#pragma unroll
    for (int ml = 0; ml < TM; ++ml) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[ml][tk * 4]), a1 = *reinterpret_cast<const float4*>(&As[ml][64 + tk * 4]);
      const float4 d0 = *reinterpret_cast<const float4*>(&Ds[ml][tn * 4]), d1 = *reinterpret_cast<const float4*>(&Ds[ml][64 + tn * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], d[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int kk = k0 + (i >> 2) * 64 + tk * 4 + (i & 3);
    if (kk >= K) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j >> 2) * 64 + tn * 4 + (j & 3);
      if (n < Cout) atomicAdd(&dW[(size_t)kk * cout_pad + n], acc[i][j]);
    }
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

__global__ void adam_kernel(float* __restrict__ P, const float* __restrict__ G, float* __restrict__ M1, float* __restrict__ M2, size_t n,
                            float lr_t, float rescale, float b1, float b2, float eps) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float g = G[i] * rescale;
    const float m = b1 * M1[i] + (1.f - b1) * g;
    const float v = b2 * M2[i] + (1.f - b2) * g * g;
    M1[i] = m; M2[i] = v;
    P[i] -= lr_t * m / (sqrtf(v) + eps);
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

void train_release(yolo_handle* h) {
  if (!h || !h->train) return;
  if (h->train->arena) cudaFree(h->train->arena);
  delete h->train;
  h->train = nullptr;
}

static float* act_ptr(const yolo_handle* h, const View& v, const void* input) {
  if (v.buf >= 0) return reinterpret_cast<float*>(h->ws + h->bufs[v.buf].offset);
  return const_cast<float*>(static_cast<const float*>(input));
}
static float* grad_ptr(const yolo_handle* h, const View& v) {
  return reinterpret_cast<float*>(h->train->arena + h->train->gbuf_off[v.buf]);
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

extern "C" int yolo_train_init(yolo_handle* h, float* params_flat, float* grads_flat, float* adam_m, float* adam_v, size_t n_flat, void* stream) {
  if (!h || !params_flat || !grads_flat || !adam_m || !adam_v) return fail(YOLO_E_BADARG, "train_init: null argument");
  if (!h->finalized || !h->ws) return hfail(h, fail(YOLO_E_STATE, "train_init: parameters must be finalized and the workspace set"));
  if (h->spec.precision != YOLO_PREC_FP32) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train_init: training runs in YOLO_PREC_FP32"));
  if (h->spec.net_type != YOLO_NET_CARNET) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train_init: CARNET only"));
  if (n_flat != yolo_train_flat_size(h)) return hfail(h, fail(YOLO_E_SHAPE, "train_init: flat buffers must hold %zu floats", yolo_train_flat_size(h)));
  for (const Op& op : h->ops)
    if (op.kind != OP_CONV || op.p_prebn >= 0 || op.out_nchw) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train_init: unsupported op in the plan"));
  YB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  train_release(h);
  TrainState* T = new TrainState();
  h->train = T;
  T->P = params_flat; T->G = grads_flat; T->M1 = adam_m; T->M2 = adam_v; T->n_flat = n_flat;
  const int B = h->spec.max_batch;
  // ---- arena layout ----
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  struct Tmp { size_t z, rmean, rvar, sums, mean, rstd, wT; };
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
    L.cin_pad = (op.in.C + 3) & ~3;
    const bool user_out = op.out.buf < -1;
    tmp[i].z = user_out ? (size_t)-1 : take((size_t)B * L.Ho * L.Wo * op.cout * 4);
    tmp[i].rmean = take(op.cout * 4); tmp[i].rvar = take(op.cout * 4);
    tmp[i].sums = take((size_t)4 * op.cout * 8);
    tmp[i].mean = take(op.cout * 4); tmp[i].rstd = take(op.cout * 4);
    tmp[i].wT = op.in.buf >= 0 ? take((size_t)op.kh * op.kw * op.cout * L.cin_pad * 4) : (size_t)-1;
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
  for (size_t i = 0; i < h->bufs.size(); ++i) T->gbuf_off[i] = take(h->bufs[i].bytes_per_image * (size_t)B);
  T->grads_bytes = off - T->grads_begin;
  const size_t lsb = yolo_loss_scratch_bytes(B, 16);
  const size_t o_ls = take(lsb);
  if (cudaMalloc(reinterpret_cast<void**>(&T->arena), off) != cudaSuccess) {
    cudaGetLastError();
    delete T; h->train = nullptr;
    return hfail(h, fail(YOLO_E_OOM, "train_init: cudaMalloc(%zu) failed", off));
  }
  T->arena_bytes = off;
  T->loss_scratch = T->arena + o_ls; T->loss_scratch_bytes = lsb;
  YB_CUDA(cudaMemsetAsync(T->arena, 0, off, st));
  for (size_t i = 0; i < h->outputs.size(); ++i) T->dheads[i] = reinterpret_cast<float*>(T->arena + dhead_off[i]);
  // ---- fill the flat parameter buffer from the loaded parameters; running stats; head z buffers = head outputs ----
  std::vector<float> hostP(n_flat, 0.f);
  for (size_t i = 0; i < h->ops.size(); ++i) {
    const Op& op = h->ops[i];
    TrainLayer& L = T->layers[i];
    const int cin = op.in.C, cout = op.cout;
    const float* Wsrc = h->params[op.p_weight].host.data();
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int r = 0; r < op.kh; ++r)
          for (int s2 = 0; s2 < op.kw; ++s2)
            hostP[L.o_w + ((size_t)(r * op.kw + s2) * cin + c) * op.cout_pad + o] = Wsrc[(((size_t)o * cin + c) * op.kh + r) * op.kw + s2];
    if (L.has_bn) {
      memcpy(&hostP[L.o_gamma], h->params[op.p_bn].host.data(), cout * 4);
      memcpy(&hostP[L.o_beta], h->params[op.p_bn + 1].host.data(), cout * 4);
    }
    if (L.has_bias) memcpy(&hostP[L.o_bias], h->params[op.p_bias].host.data(), cout * 4);
    L.rmean = reinterpret_cast<float*>(T->arena + tmp[i].rmean); L.rvar = reinterpret_cast<float*>(T->arena + tmp[i].rvar);
    L.sums = reinterpret_cast<double*>(T->arena + tmp[i].sums);
    L.mean = reinterpret_cast<float*>(T->arena + tmp[i].mean); L.rstd = reinterpret_cast<float*>(T->arena + tmp[i].rstd);
    L.wT = tmp[i].wT == (size_t)-1 ? nullptr : reinterpret_cast<float*>(T->arena + tmp[i].wT);
    if (op.out.buf < -1) L.z = reinterpret_cast<float*>(T->arena + head_off[-2 - op.out.buf]);
    else L.z = reinterpret_cast<float*>(T->arena + tmp[i].z);
    if (L.has_bn) {
      YB_CUDA(cudaMemcpyAsync(L.rmean, h->params[op.p_bn + 2].host.data(), cout * 4, cudaMemcpyHostToDevice, st));
      YB_CUDA(cudaMemcpyAsync(L.rvar, h->params[op.p_bn + 3].host.data(), cout * 4, cudaMemcpyHostToDevice, st));
    }
  }
  YB_CUDA(cudaMemcpyAsync(T->P, hostP.data(), n_flat * 4, cudaMemcpyHostToDevice, st));
  YB_CUDA(cudaMemsetAsync(T->M1, 0, n_flat * 4, st));
  YB_CUDA(cudaMemsetAsync(T->M2, 0, n_flat * 4, st));
  YB_CUDA(cudaMemsetAsync(T->G, 0, n_flat * 4, st));
  YB_CUDA(cudaStreamSynchronize(st));
  // inference and training share the weights from now on
  for (size_t i = 0; i < h->ops.size(); ++i) h->ops[i].w_f32 = T->P + T->layers[i].o_w;
  return YOLO_OK;
}

// forward (train-mode BN) + targets/losses + backward: fills the flat gradient buffer (d sum(losses) / d param)
extern "C" int yolo_train_forward_backward(yolo_handle* h, const void* input, int in_layout, const float* labels, int batch, int n_obj,
                                           const yolo_loss_params* lp, float* out_losses, void* stream) {
  if (!h || !h->train) return fail(YOLO_E_STATE, "train: yolo_train_init has not been called");
  if (!input || !labels || !lp || !out_losses) return hfail(h, fail(YOLO_E_BADARG, "train: null argument"));
  if (batch < 1 || batch > h->spec.max_batch) return hfail(h, fail(YOLO_E_SHAPE, "train: batch=%d outside [1,%d]", batch, h->spec.max_batch));
  YB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  TrainState* T = h->train;
  const int launches0 = g_launches;
  const int lay_in = in_layout == YOLO_IN_NCHW_F32 ? 1 : 2;
  // ---------------- forward ----------------
  for (size_t i = 0; i < h->ops.size(); ++i) {
    const Op& op = h->ops[i];
    TrainLayer& L = T->layers[i];
    const int M = batch * L.Ho * L.Wo, C = op.cout;
    ConvDesc d;
    memset(&d, 0, sizeof(d));
    d.in = act_ptr(h, op.in, input); d.in_dtype = DT_F32; d.N = batch; d.H = op.in.H; d.W = op.in.W; d.Cin = op.in.C;
    d.in_cpitch = op.in.cpitch; d.in_coff = op.in.coff;
    d.kh = op.kh; d.kw = op.kw; d.stride = op.stride; d.pad = op.pad; d.Cout = C;
    d.w_f32 = T->P + L.o_w; d.cout_pad = op.cout_pad;
    d.shift = L.has_bias && !L.has_bn ? T->P + L.o_bias : nullptr;        // head convs: z = conv + bias is the output itself
    d.act = ACT_NONE;
    d.out = L.z; d.out_dtype = DT_F32; d.Ho = L.Ho; d.Wo = L.Wo; d.out_cpitch = C; d.out_coff = 0;
    int rc = launch_conv_simt(d, op.in.buf == -1 ? lay_in : 0, st);
    if (rc) return hfail(h, rc);
    if (!L.has_bn) continue;
    YB_CUDA(cudaMemsetAsync(L.sums, 0, (size_t)4 * C * 8, st));
    dim3 g((C + 31) / 32, std::max(1, std::min(M / 64, 512)));
    channel_sums_kernel<0><<<g, 256, 0, st>>>(L.z, M, C, L.sums, nullptr, 0, 0, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, 0);
    bn_finalize_fwd_kernel<<<(C + 127) / 128, 128, 0, st>>>(L.sums, M, C, L.mean, L.rstd, L.rmean, L.rvar);
    const float* res = op.has_res ? act_ptr(h, op.res, input) : nullptr;
    bn_act_fwd_kernel<<<grid_for((size_t)M * C), 256, 0, st>>>(L.z, M, C, L.mean, L.rstd, T->P + L.o_gamma, T->P + L.o_beta, op.act, res,
                                                               op.res.cpitch, op.res.coff, act_ptr(h, op.out, input), op.out.cpitch, op.out.coff,
                                                               op.upsample2, L.Ho, L.Wo);
    g_launches += 3;
  }
  YB_CUDA(cudaGetLastError());
  // ---------------- targets, losses, d loss / d heads ----------------
  yolo_decode_geom g;
  memset(&g, 0, sizeof(g));
  const yolo_spec& s = h->spec;
  g.height = s.height; g.width = s.width; g.n_scales = s.n_scales; g.n_anchors = s.n_anchors; g.channels_per_anchor = s.channels_per_anchor;
  for (int i = 0; i < s.n_scales; ++i) {
    g.step[i] = 1 << (s.n_layers - s.n_scales + 1 + i);
    for (int a = 0; a < s.n_anchors; ++a) { g.anchors[i][a][0] = s.anchors[i][a][0]; g.anchors[i][a][1] = s.anchors[i][a][1]; }
  }
  const void* heads[YOLO_MAX_SCALES];
  void* dheads[YOLO_MAX_SCALES];
  for (size_t i = 0; i < h->ops.size(); ++i)
    if (h->ops[i].out.buf < -1) heads[-2 - h->ops[i].out.buf] = T->layers[i].z;
  for (int i = 0; i < s.n_scales; ++i) dheads[i] = T->dheads[i];
  if (n_obj > 16) return hfail(h, fail(YOLO_E_UNSUPPORTED, "train: at most 16 labels per image"));
  int rc = yolo_loss_targets(&g, heads, labels, batch, n_obj, lp, T->loss_scratch, out_losses, dheads, nullptr, stream);
  if (rc) return hfail(h, rc);
  // ---------------- backward ----------------
  YB_CUDA(cudaMemsetAsync(T->arena + T->grads_begin, 0, T->grads_bytes, st));
  YB_CUDA(cudaMemsetAsync(T->G, 0, T->n_flat * 4, st));
  for (int i = (int)h->ops.size() - 1; i >= 0; --i) {
    const Op& op = h->ops[i];
    TrainLayer& L = T->layers[i];
    const int M = batch * L.Ho * L.Wo, C = op.cout;
    float* dz;
    YB_CUDA(cudaMemsetAsync(L.sums, 0, (size_t)4 * C * 8, st));
    dim3 gs((C + 31) / 32, std::max(1, std::min(M / 64, 512)));
    if (L.has_bn) {
      const float* dy = grad_ptr(h, op.out);
      float* dres = op.has_res ? grad_ptr(h, op.res) : nullptr;
      channel_sums_kernel<1><<<gs, 256, 0, st>>>(L.z, M, C, L.sums, dy, op.out.cpitch, op.out.coff, op.upsample2, L.Ho, L.Wo, L.mean, L.rstd,
                                                 T->P + L.o_gamma, T->P + L.o_beta, op.act, dres, op.res.cpitch, op.res.coff);
      bn_bwd_apply_kernel<<<grid_for((size_t)M * C), 256, 0, st>>>(L.z, M, C, L.sums, dy, op.out.cpitch, op.out.coff, op.upsample2, L.Ho, L.Wo,
                                                                   L.mean, L.rstd, T->P + L.o_gamma, T->P + L.o_beta, op.act);
      bn_param_grads_kernel<<<(C + 127) / 128, 128, 0, st>>>(L.sums, C, T->G + L.o_gamma, T->G + L.o_beta);
      dz = L.z;
      g_launches += 3;
    } else {
      dz = T->dheads[-2 - op.out.buf];                      // head conv: dz = d loss / d head
      channel_sums_kernel<0><<<gs, 256, 0, st>>>(dz, M, C, L.sums, nullptr, 0, 0, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, 0);
      bias_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(L.sums, C, T->G + L.o_bias);
      g_launches += 2;
    }
    // weight gradient
    const int K = op.kh * op.kw * op.in.C;
    int slabs = std::max(1, std::min(M / 512, 64));
    dim3 gw((K + 63) / 64, (C + 63) / 64, slabs);
    const float* xin = act_ptr(h, op.in, input);
    if (op.in.buf != -1 && op.in.C % 4 == 0 && op.in.cpitch % 4 == 0 && op.in.coff % 4 == 0 && C % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(xin) & 15) == 0 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0) {
      const int tiles = ((K + 127) / 128) * ((C + 127) / 128);
      int slabs2 = std::max(1, std::min(M / 256, std::max(1, 592 / tiles)));          // ~4 CTAs per SM in total
      dim3 gw2((K + 127) / 128, (C + 127) / 128, slabs2);
      wgrad128_kernel<<<gw2, 256, 0, st>>>(xin, batch, op.in.H, op.in.W, op.in.C, op.in.cpitch, op.in.coff, dz, L.Ho, L.Wo, C, op.kh, op.kw,
                                           op.stride, op.pad, T->G + L.o_w, op.cout_pad);
    } else
    wgrad_kernel<<<gw, 256, 0, st>>>(xin, batch, op.in.H, op.in.W, op.in.C, op.in.cpitch, op.in.coff,
                                     op.in.buf == -1 ? lay_in : 0, dz, L.Ho, L.Wo, C, op.kh, op.kw, op.stride, op.pad, T->G + L.o_w, op.cout_pad);
    ++g_launches;
    // data gradient (not needed for the network input)
    if (op.in.buf < 0) continue;
    wflip_kernel<<<grid_for((size_t)K * C), 256, 0, st>>>(T->P + L.o_w, op.kh, op.kw, op.in.C, C, op.cout_pad, L.wT, L.cin_pad);
    ++g_launches;
    ConvDesc d;
    memset(&d, 0, sizeof(d));
    d.in = dz; d.in_dtype = DT_F32; d.N = batch; d.H = L.Ho; d.W = L.Wo; d.Cin = C; d.in_cpitch = C; d.in_coff = 0;
    d.kh = op.kh; d.kw = op.kw; d.stride = 1; d.pad = op.kh - 1 - op.pad; d.in_dil = op.stride; d.Cout = op.in.C;
    d.w_f32 = L.wT; d.cout_pad = L.cin_pad;
    d.act = ACT_NONE;
    float* dx = grad_ptr(h, op.in);
    d.res = dx; d.res_cpitch = op.in.cpitch; d.res_coff = op.in.coff;        // accumulate: several consumers may feed one tensor
    d.out = dx; d.out_dtype = DT_F32; d.Ho = op.in.H; d.Wo = op.in.W; d.out_cpitch = op.in.cpitch; d.out_coff = op.in.coff;
    rc = launch_conv_simt(d, 0, st);
    if (rc) return hfail(h, rc);
  }
  YB_CUDA(cudaGetLastError());
  h->last_launches = g_launches - launches0;
  return YOLO_OK;
}

// trainer.step(batch_size): G is expected to hold the SUM over ranks (all-reduce done by the caller); rescale = 1/batch_size
extern "C" int yolo_train_apply(yolo_handle* h, float lr, float beta1, float beta2, float eps, float rescale_grad, void* stream) {
  if (!h || !h->train) return fail(YOLO_E_STATE, "train_apply: yolo_train_init has not been called");
  YB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  TrainState* T = h->train;
  const int t = ++T->step_count;
  const float lr_t = lr * sqrtf(1.f - powf(beta2, (float)t)) / (1.f - powf(beta1, (float)t));     // mxnet.optimizer.Adam
  adam_kernel<<<grid_for(T->n_flat), 256, 0, st>>>(T->P, T->G, T->M1, T->M2, T->n_flat, lr_t, rescale_grad, beta1, beta2, eps);
  ++g_launches;
  for (size_t i = 0; i < h->ops.size(); ++i) {          // keep the inference epilogues consistent with the trained parameters
    Op& op = h->ops[i];
    TrainLayer& L = T->layers[i];
    if (!op.scale) continue;
    refold_kernel<<<(op.cout + 127) / 128, 128, 0, st>>>(L.has_bn ? T->P + L.o_gamma : nullptr, L.has_bn ? T->P + L.o_beta : nullptr, L.rmean, L.rvar,
                                                         L.has_bias ? T->P + L.o_bias : nullptr, op.cout, op.scale, op.shift);
  }
  YB_CUDA(cudaGetLastError());
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
