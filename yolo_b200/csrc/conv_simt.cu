// fp32-accurate implicit-GEMM convolution on the CUDA cores (FFMA), NHWC, with the fused
// prologue/epilogue set the reference's blocks need.  This is the YOLO_PREC_FP32 (parity-grade) path
// and the fallback for shapes the tcgen05 kernel does not take (Cin % 64 != 0: stems, DenseNet).
//
// Replaces the MXNet operator chains (reference, file:line):
//   gluoncv _conv2d = Convolution(no bias) -> BatchNorm(eps 1e-5) -> LeakyReLU(0.1)    yolo_modules/basic_yolo.py:20-26,118-121
//   DarknetBasicBlockV3 residual add (elemwise_add)                                     yolo_modules/basic_yolo.py:26
//   _upsample (repeat x2) + concat(dim=1)                                               car/utils.py:92-93
//   YOLOOutput: 1x1 Convolution + bias, transpose(0,2,3,1)                              yolo_modules/basic_yolo.py:98-103
//   DenseNet BN -> ReLU -> conv (pre-activation), concat([x, new])                      licence_plate/LP_detection.py:70-93
//   MaxPool 3/2 p1, AvgPool 2/2                                                         licence_plate/LP_detection.py:74
//   cv_img_2_ndarray (/255, HWC->CHW) fused into the stem's gather                      yolo_modules/yolo_gluon.py:335-357
//
// GEMM view: M = N*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin with k = (r*kw + s)*Cin + c.
// CTA tile 128(M) x 64(N) x 16(K), 256 threads, 8x4 accumulators per thread, register-staged
// double buffering of the shared-memory tiles.
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.cuh"

namespace yb {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int AS_PITCH = BM + 4;

enum InLayout : int { IN_NHWC = 0, IN_NCHW_F32 = 1, IN_NHWC_U8 = 2 };

struct ConvKArgs {
  ConvDesc d;
  int in_layout;
  int M, K;
};

__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ void load8(const float* p, float v[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float v[8]) {
  uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

__device__ __forceinline__ void store4(float* p, const float v[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float v[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<unsigned*>(&a);
  u.y = *reinterpret_cast<unsigned*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void store1(float* p, float v) { *p = v; }
__device__ __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void load4(const float* p, float v[4]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float v[4]) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

__device__ __forceinline__ float ld_as_float(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ void load8(const __half* p, float v[8]) {
  uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load4(const __half* p, float v[4]) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&u);
  float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(__half* p, const float v[4]) {
  __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<unsigned*>(&a);
  u.y = *reinterpret_cast<unsigned*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void store1(__half* p, float v) { *p = __float2half_rn(v); }

// ---- storage formats ---------------------------------------------------------------------------------------
// F32 / BF16: one plane.  BF16X3: v = p0 + p1 + p2 (exact 24-bit split).  F16X2: v = p0 + p1 (22-bit fp16 split, p1 unscaled);
// plane q lives at element offset q * plane_stride.
struct FmtF32 { using T = float; static constexpr int NP = 1; static constexpr int ALIGN = 4; };
struct FmtBF16 { using T = __nv_bfloat16; static constexpr int NP = 1; static constexpr int ALIGN = 8; };
struct FmtBF16X3 { using T = __nv_bfloat16; static constexpr int NP = 3; static constexpr int ALIGN = 8; };
struct FmtF16X2 { using T = __half; static constexpr int NP = 2; static constexpr int ALIGN = 8; };

template <class F> __device__ __forceinline__ float plane_weight(int q) {
  return (F::NP == 2 && q == 1) ? kF16LoScaleInv : 1.f;
}
template <class F>
__device__ __forceinline__ void load8f(const typename F::T* p, long long ps, float v[8]) {
  load8(p, v);
#pragma unroll
  for (int q = 1; q < F::NP; ++q) {
    float t[8];
    load8(p + q * ps, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(t[i], plane_weight<F>(q), v[i]);
  }
}
template <class F>
__device__ __forceinline__ void load4f(const typename F::T* p, long long ps, float v[4]) {
  load4(p, v);
#pragma unroll
  for (int q = 1; q < F::NP; ++q) {
    float t[4];
    load4(p + q * ps, t);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = fmaf(t[i], plane_weight<F>(q), v[i]);
  }
}
template <class F>
__device__ __forceinline__ float ld1f(const typename F::T* p, long long ps) {
  float v = ld_as_float(p);
#pragma unroll
  for (int q = 1; q < F::NP; ++q) v = fmaf(ld_as_float(p + q * ps), plane_weight<F>(q), v);
  return v;
}
// split one value into the planes of format F (returned as floats that are exactly representable in F::T)
template <class F>
__device__ __forceinline__ void split_planes(float v, float h[F::NP], int& sat) {
  if (F::NP == 1) { h[0] = v; return; }
  if (F::NP == 3) {
#pragma unroll
    for (int q = 0; q < 3; ++q) { h[q] = __bfloat162float(__float2bfloat16_rn(v)); v -= h[q]; }
    return;
  }
  if (fabsf(v) > kF16Max) sat = 1;                       // the high plane saturates: flagged (ConvDesc::sat_flag), never silent
  float c = fminf(fmaxf(v, -kF16Max), kF16Max);
  h[0] = __half2float(__float2half_rn(c));
  h[1 % F::NP] = fminf(fmaxf((v - h[0]) * kF16LoScale, -kF16Max), kF16Max);
}
template <class F>
__device__ __forceinline__ void store4f(typename F::T* p, long long ps, const float v[4], int& sat) {
  float h[4][F::NP];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_planes<F>(v[i], h[i], sat);
#pragma unroll
  for (int q = 0; q < F::NP; ++q) {
    const float t[4] = {h[0][q], h[1][q], h[2][q], h[3][q]};
    store4(p + q * ps, t);
  }
}
template <class F>
__device__ __forceinline__ void store1f(typename F::T* p, long long ps, float v, int& sat) {
  float h[F::NP];
  split_planes<F>(v, h, sat);
#pragma unroll
  for (int q = 0; q < F::NP; ++q) store1(p + q * ps, h[q]);
}

template <class FI, class FO, bool VEC>
__global__ void __launch_bounds__(NT) conv_simt_kernel(const __grid_constant__ ConvKArgs args) {
  using TIn = typename FI::T;
  using TOut = typename FO::T;
  const ConvDesc& d = args.d;
  __shared__ __align__(16) float As[2][BK][AS_PITCH];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int M = args.M, K = args.K;
  const int HoWo = d.Ho * d.Wo;

  // ---- A-load role: one pixel, 8 consecutive k per thread --------------------------------------
  const int am = tid >> 1, akh = (tid & 1) * 8;
  const int a_m = m0 + am;
  const bool a_valid = a_m < M;
  int a_n = 0, a_ih0 = 0, a_iw0 = 0;
  if (a_valid) {
    a_n = a_m / HoWo;
    int rem = a_m - a_n * HoWo;
    int oh = rem / d.Wo, ow = rem - oh * d.Wo;
    a_ih0 = oh * d.stride - d.pad;
    a_iw0 = ow * d.stride - d.pad;
  }
  // ---- B-load role: one k row, 4 consecutive couts ------------------------------------------------
  const int bk = tid >> 4, bn4 = (tid & 15) * 4;
  const bool b_ncol_ok = (n0 + bn4) < d.cout_pad;

  float ra[8];
  float4 rb;

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) ra[i] = 0.f;
    if (VEC) {
      int kg = k0 + akh;                       // Cin % 16 == 0: the 16-k tile sits inside one tap
      int tap = kg / d.Cin, c = kg - tap * d.Cin;
      int r = tap / d.kw, s = tap - r * d.kw;
      int ih = a_ih0 + r, iw = a_iw0 + s;
      bool ok = a_valid;
      if (d.in_dil > 1) {                      // transposed conv: only every in_dil-th position of the dilated input is real
        ok = ok && ih >= 0 && iw >= 0 && (ih % d.in_dil) == 0 && (iw % d.in_dil) == 0;
        ih /= d.in_dil; iw /= d.in_dil;
      }
      if (ok && (unsigned)ih < (unsigned)d.H && (unsigned)iw < (unsigned)d.W) {
        const TIn* p = static_cast<const TIn*>(d.in) + ((size_t)(a_n * d.H + ih) * d.W + iw) * d.in_cpitch + d.in_coff + c;
        load8f<FI>(p, d.in_plane_stride, ra);
        if (d.pre_scale) {
#pragma unroll
          for (int i = 0; i < 8; ++i) ra[i] = fmaxf(fmaf(ra[i], __ldg(d.pre_scale + c + i), __ldg(d.pre_shift + c + i)), 0.f);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int kg = k0 + akh + i;
        if (a_valid && kg < K) {
          int tap = kg / d.Cin, c = kg - tap * d.Cin;
          int r = tap / d.kw, s = tap - r * d.kw;
          int ih = a_ih0 + r, iw = a_iw0 + s;
          bool ok = true;
          if (d.in_dil > 1) {
            ok = ih >= 0 && iw >= 0 && (ih % d.in_dil) == 0 && (iw % d.in_dil) == 0;
            ih /= d.in_dil; iw /= d.in_dil;
          }
          if (ok && (unsigned)ih < (unsigned)d.H && (unsigned)iw < (unsigned)d.W) {
            float v;
            if (args.in_layout == IN_NCHW_F32) {
              v = __ldg(static_cast<const float*>(d.in) + ((size_t)(a_n * d.Cin + c) * d.H + ih) * d.W + iw);
            } else if (args.in_layout == IN_NHWC_U8) {
              v = (float)__ldg(static_cast<const unsigned char*>(d.in) + ((size_t)(a_n * d.H + ih) * d.W + iw) * d.Cin + c) / 255.f;
            } else {
              v = ld1f<FI>(static_cast<const TIn*>(d.in) + ((size_t)(a_n * d.H + ih) * d.W + iw) * d.in_cpitch + d.in_coff + c, d.in_plane_stride);
            }
            if (d.pre_scale) v = fmaxf(fmaf(v, __ldg(d.pre_scale + c), __ldg(d.pre_shift + c)), 0.f);
            ra[i] = v;
          }
        }
      }
    }
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b_ncol_ok && (k0 + bk) < K)
      rb = __ldg(reinterpret_cast<const float4*>(d.w_f32 + (size_t)(k0 + bk) * d.cout_pad + n0 + bn4));
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[buf][akh + i][am] = ra[i];
    *reinterpret_cast<float4*>(&Bs[buf][bk][bn4]) = rb;
  };

  const int tm = tid >> 4, tn = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue -----------------------------------------------------------------------------------
  const int nb = n0 + tn * 4;
  if (nb >= d.Cout) return;
  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bool ok = nb + j < d.Cout;
    sc[j] = ok && d.scale ? __ldg(d.scale + nb + j) : 1.f;
    sh[j] = ok && d.shift ? __ldg(d.shift + nb + j) : 0.f;
  }
  const float dyn = d.dyn_scale ? __ldg(d.dyn_scale) : 1.f;
  const bool full4 = nb + 3 < d.Cout;
  const bool vec_out = full4 && !d.out_nchw && ((d.out_cpitch | d.out_coff) & 3) == 0;
  const bool vec_res = full4 && d.res && ((d.res_cpitch | d.res_coff) & 3) == 0;
  TOut* out = static_cast<TOut*>(d.out);
  const TOut* res = static_cast<const TOut*>(d.res);
  int sat = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + tm * 8 + i;
    if (m >= M) break;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float y = fmaf(acc[i][j] * dyn, sc[j], sh[j]);
      if (d.act == ACT_LEAKY) y = y > 0.f ? y : 0.1f * y;
      else if (d.act == ACT_RELU) y = fmaxf(y, 0.f);
      v[j] = y;
    }
    if (res) {
      const TOut* rp = res + (size_t)m * d.res_cpitch + d.res_coff + nb;
      if (vec_res) {
        float r4[4];
        load4f<FO>(rp, d.res_plane_stride, r4);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += r4[j];
      } else {
        for (int j = 0; j < 4 && nb + j < d.Cout; ++j) v[j] += ld1f<FO>(rp + j, d.res_plane_stride);
      }
    }
    if (d.out_nchw) {
      int n = m / HoWo, rem = m - n * HoWo;
      for (int j = 0; j < 4 && nb + j < d.Cout; ++j)
        store1f<FO>(out + ((size_t)n * d.Cout + nb + j) * HoWo + rem, d.out_plane_stride, v[j], sat);
    } else if (d.upsample2) {
      int n = m / HoWo, rem = m - n * HoWo;
      int oh = rem / d.Wo, ow = rem - oh * d.Wo;
      const int W2 = 2 * d.Wo;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        size_t pix = ((size_t)n * 2 * d.Ho + 2 * oh + (q >> 1)) * W2 + 2 * ow + (q & 1);
        TOut* op = out + pix * d.out_cpitch + d.out_coff + nb;
        if (vec_out) store4f<FO>(op, d.out_plane_stride, v, sat);
        else for (int j = 0; j < 4 && nb + j < d.Cout; ++j) store1f<FO>(op + j, d.out_plane_stride, v[j], sat);
      }
    } else {
      TOut* op = out + (size_t)m * d.out_cpitch + d.out_coff + nb;
      if (vec_out) store4f<FO>(op, d.out_plane_stride, v, sat);
      else for (int j = 0; j < 4 && nb + j < d.Cout; ++j) store1f<FO>(op + j, d.out_plane_stride, v[j], sat);
    }
  }
  if (sat && d.sat_flag) atomicOr(d.sat_flag, YOLO_SAT_ACT_FFMA);
}

template <class FI, class FO>
static void launch_t(const ConvKArgs& a, bool vec, dim3 grid, cudaStream_t st) {
  if (vec) conv_simt_kernel<FI, FO, true><<<grid, NT, 0, st>>>(a);
  else conv_simt_kernel<FI, FO, false><<<grid, NT, 0, st>>>(a);
}

int launch_conv_simt(const ConvDesc& d, int in_layout, cudaStream_t st) {
  ConvKArgs a;
  a.d = d;
  a.in_layout = in_layout;
  a.M = d.N * d.Ho * d.Wo;
  a.K = d.kh * d.kw * d.Cin;
  if (a.M <= 0 || d.Cout <= 0) return fail(YOLO_E_SHAPE, "conv: empty problem M=%d Cout=%d", a.M, d.Cout);
  if (d.out_nchw && d.out_dtype != DT_F32) return fail(YOLO_E_BADARG, "conv: NCHW output is fp32 only");
  const int idt = in_layout == IN_NHWC ? d.in_dtype : DT_F32;
  const int in_align = idt == DT_F32 ? 4 : 8;
  bool vec = in_layout == IN_NHWC && d.Cin % 16 == 0 && d.in_cpitch % in_align == 0 && d.in_coff % in_align == 0 &&
             (reinterpret_cast<uintptr_t>(d.in) & 15) == 0 && d.in_plane_stride % 8 == 0;
  dim3 grid((a.M + BM - 1) / BM, (d.Cout + BN - 1) / BN);
  const int odt = d.out_dtype;
  if (idt == DT_F32 && odt == DT_F32) launch_t<FmtF32, FmtF32>(a, vec, grid, st);
  else if (idt == DT_F32 && odt == DT_BF16) launch_t<FmtF32, FmtBF16>(a, vec, grid, st);
  else if (idt == DT_F32 && odt == DT_BF16X3) launch_t<FmtF32, FmtBF16X3>(a, vec, grid, st);
  else if (idt == DT_F32 && odt == DT_F16X2) launch_t<FmtF32, FmtF16X2>(a, vec, grid, st);
  else if (idt == DT_BF16 && odt == DT_BF16) launch_t<FmtBF16, FmtBF16>(a, vec, grid, st);
  else if (idt == DT_BF16 && odt == DT_F32) launch_t<FmtBF16, FmtF32>(a, vec, grid, st);
  else if (idt == DT_BF16X3 && odt == DT_BF16X3) launch_t<FmtBF16X3, FmtBF16X3>(a, vec, grid, st);
  else if (idt == DT_BF16X3 && odt == DT_F32) launch_t<FmtBF16X3, FmtF32>(a, vec, grid, st);
  else if (idt == DT_F16X2 && odt == DT_F16X2) launch_t<FmtF16X2, FmtF16X2>(a, vec, grid, st);
  else if (idt == DT_F16X2 && odt == DT_F32) launch_t<FmtF16X2, FmtF32>(a, vec, grid, st);
  else return fail(YOLO_E_UNSUPPORTED, "conv: dtype combination in=%d out=%d", idt, odt);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

// ---- stem: 3x3 convolution on the 3-channel network input ----------------------------------------------
// One thread per output pixel, the 27 input taps in registers, weights broadcast from shared memory, 16 output
// channels at a time.  Reads the reference's input contract directly - NCHW fp32 in [0,1] (cv_img_2_ndarray,
// yolo_modules/yolo_gluon.py:335-357) or the uint8 HWC camera frame with the /255 fused - and writes the first NHWC
// activation in the precision's storage format.  (K = 27 is too small for the tensor cores; this layer is
// bound by its output write.)
constexpr int STEM_CO = 16;
// PX = output pixels per thread (consecutive in the flat pixel index).  With one pixel per thread the kernel is co-limited by the FMA
// pipe and the shared-memory pipe (one broadcast LDS.128 of weights per 4 FMAs); two pixels per thread reuse every weight load twice.
template <class FO, int LAYOUT, int PX>
__global__ void __launch_bounds__(256) stem3x3_kernel(const __grid_constant__ ConvKArgs args) {
  using TOut = typename FO::T;
  const ConvDesc& d = args.d;
  extern __shared__ float s_w[];                 // [27][cout_pad] then scale[Cout], shift[Cout]
  const int cp = d.cout_pad;
  for (int i = threadIdx.x; i < 27 * cp; i += blockDim.x) s_w[i] = __ldg(d.w_f32 + i);
  float* s_sc = s_w + 27 * cp;
  float* s_sh = s_sc + d.Cout;
  for (int i = threadIdx.x; i < d.Cout; i += blockDim.x) {
    s_sc[i] = d.scale ? __ldg(d.scale + i) : 1.f;
    s_sh[i] = d.shift ? __ldg(d.shift + i) : 0.f;
  }
  __syncthreads();
  const int HoWo = d.Ho * d.Wo;
  const int m_blk0 = blockIdx.x * blockDim.x * PX;
  float x[PX][27];
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int m = m_blk0 + threadIdx.x * PX + p;
    const int mm = m < args.M ? m : args.M - 1;
    const int n = mm / HoWo, rem = mm - n * HoWo;
    const int oh = rem / d.Wo, ow = rem - oh * d.Wo;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int ih = oh * d.stride - d.pad + r, iw = ow * d.stride - d.pad + q;
        const bool ok = (unsigned)ih < (unsigned)d.H && (unsigned)iw < (unsigned)d.W;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = 0.f;
          if (ok) {
            if (LAYOUT == IN_NCHW_F32) v = __ldg(static_cast<const float*>(d.in) + ((size_t)(n * 3 + c) * d.H + ih) * d.W + iw);
            else v = (float)__ldg(static_cast<const unsigned char*>(d.in) + ((size_t)(n * d.H + ih) * d.W + iw) * 3 + c) / 255.f;
          }
          x[p][(r * 3 + q) * 3 + c] = v;
        }
      }
  }
  // Results are staged in shared memory as whole pixel rows so that the block writes its 256*PX consecutive pixels with
  // fully coalesced 16-byte stores (a thread-per-pixel store pattern touches 32 different lines per instruction).
  TOut* s_out = reinterpret_cast<TOut*>(s_sh + d.Cout);          // [NP][256*PX][Cout]
  const int blk_pix = (int)blockDim.x * PX;
  const bool staged = args.in_layout >= 0 && (d.Cout * sizeof(TOut)) % 16 == 0 && (d.out_cpitch * sizeof(TOut)) % 16 == 0 &&
                      (d.out_coff * sizeof(TOut)) % 16 == 0 && (d.out_plane_stride * sizeof(TOut)) % 16 == 0;
  constexpr int SW_E = 16 / (int)sizeof(TOut);                     // elements per 16-byte vector
  const int sw_vpp = d.Cout / SW_E;                                // vectors per pixel row (power of two: Cout is 16 or 32)
  const int sw_rpl = max(1, 128 / (d.Cout * (int)sizeof(TOut)));   // pixel rows per 128-byte bank line
  int sat = 0;
  for (int o0 = 0; o0 < d.Cout; o0 += STEM_CO) {
    float acc[PX][STEM_CO];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int j = 0; j < STEM_CO; ++j) acc[p][j] = 0.f;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
#pragma unroll
      for (int j4 = 0; j4 < STEM_CO; j4 += 4) {
        const float4 w = *reinterpret_cast<const float4*>(&s_w[k * cp + o0 + j4]);
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          acc[p][j4] = fmaf(x[p][k], w.x, acc[p][j4]);
          acc[p][j4 + 1] = fmaf(x[p][k], w.y, acc[p][j4 + 1]);
          acc[p][j4 + 2] = fmaf(x[p][k], w.z, acc[p][j4 + 2]);
          acc[p][j4 + 3] = fmaf(x[p][k], w.w, acc[p][j4 + 3]);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      const int pl = threadIdx.x * PX + p, m = m_blk0 + pl;          // pixel within the block / flat pixel
      const int sw_x = (pl / sw_rpl) & (sw_vpp - 1);
#pragma unroll
      for (int j4 = 0; j4 < STEM_CO; j4 += 4) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float y = fmaf(acc[p][j4 + j], s_sc[o0 + j4 + j], s_sh[o0 + j4 + j]);
          if (d.act == ACT_LEAKY) y = y > 0.f ? y : 0.1f * y;
          else if (d.act == ACT_RELU) y = fmaxf(y, 0.f);
          v[j] = m < args.M ? y : 0.f;
        }
        if (staged) {
          // 16-byte vectors of a pixel row are XOR-swizzled with the pixel index: a plain [pixel][Cout] staging buffer puts the
          // 32 threads of a store on 2-4 banks (row pitch 64/128 bytes), 16-way conflicts; un-swizzled again by the copy-out
          const int e = o0 + j4, vi = e / SW_E, within = e - vi * SW_E;
          store4f<FO>(s_out + (size_t)pl * d.Cout + ((vi ^ sw_x) * SW_E) + within, (long long)blk_pix * d.Cout, v, sat);
        } else if (m < args.M) store4f<FO>(static_cast<TOut*>(d.out) + (size_t)m * d.out_cpitch + d.out_coff + o0 + j4, d.out_plane_stride, v, sat);
      }
    }
  }
  if (sat && d.sat_flag) atomicOr(d.sat_flag, YOLO_SAT_ACT_FFMA);
  if (staged) {
    __syncthreads();
    const int rows = min(blk_pix, args.M - m_blk0);
    const int vec_per_plane = rows * d.Cout * (int)sizeof(TOut) / 16;
#pragma unroll
    for (int q = 0; q < FO::NP; ++q) {
      const uint4* src = reinterpret_cast<const uint4*>(s_out + (size_t)q * blk_pix * d.Cout);
      TOut* dst = static_cast<TOut*>(d.out) + (size_t)q * d.out_plane_stride + (size_t)m_blk0 * d.out_cpitch + d.out_coff;
      for (int i = threadIdx.x; i < vec_per_plane; i += blockDim.x) {
        const int pix = i / sw_vpp, vi = i - pix * sw_vpp;
        *reinterpret_cast<uint4*>(dst + (size_t)pix * d.out_cpitch + vi * SW_E) = src[pix * sw_vpp + (vi ^ ((pix / sw_rpl) & (sw_vpp - 1)))];
      }
    }
  }
}

// ---- stem, tiled: 16 x 16*PX output pixels per block, input patch staged in shared memory, packed FFMA2 --------------------
// The pixel-per-thread kernel above gathers its 27 taps from global memory (27 byte loads + 27 IEEE divisions per pixel for camera
// frames) and issues one FFMA per multiply-add.  Here the block converts its (16+2) x (32+2) x 3 input patch ONCE into shared
// memory (coalesced row segments; /255 once per input element), each thread owns PX pixels (columns tx + 16p of one tile row)
// x all CO output channels, and the multiply-adds are `fma.rn.f32x2` on (channel pair) registers - two IEEE fp32 FMAs per issue
// slot, bit-identical to scalar fmaf in the same k order.  Results leave through the same swizzled staging as above.
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long dup2(float x) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(x));
  return d;
}
// WCONST: the weights travel as a kernel parameter (3.4 KB in constant bank 0).  With every index unrolled the compiler feeds them to
// FFMA2 through uniform registers (LDCU.128), so the shared-memory pipe - saturated at 95 % by the broadcast weight loads of the
// shared-memory variant (profiles/r2_ncu_stem.txt) - only carries the input patch and the staged results.  Training reads the master
// weights from the flat device buffer and keeps the shared-memory variant.
constexpr int STEM_TH = 16;
template <int N>
struct StemWeights { float w[N]; int co_off; };               // [KS*KS*3][CO] slice of the weight matrix for output channels [co_off, co_off + CO)
template <class FO, int LAYOUT, int CO, int PX, bool WCONST, int KS, int STRIDE>
__global__ void __launch_bounds__(256, PX == 2 ? 2 : 4)
stem_tile_kernel(const __grid_constant__ ConvKArgs args, const __grid_constant__ StemWeights<WCONST ? KS * KS * 3 * CO : 1> wp) {
  using TOut = typename FO::T;
  constexpr int TW = 16 * PX, TAPS = KS * KS * 3;
  constexpr int PH = (STEM_TH - 1) * STRIDE + KS, PW = (TW - 1) * STRIDE + KS, IW = PW * 3;
  constexpr int IPITCH = ((IW + 15) / 32) * 32 + 16;               // >= IW and = 16 mod 32: the two tile rows of a warp hit disjoint banks (stride 1)
  const ConvDesc& d = args.d;
  extern __shared__ __align__(16) float s_w[];                     // [TAPS][CO] (shared-memory variant), scale[CO], shift[CO], patch[PH][IPITCH], staged out
  float* s_sc = s_w + (WCONST ? 0 : TAPS * CO);
  float* s_sh = s_sc + CO;
  float* s_in = s_sh + CO;
  TOut* s_out = reinterpret_cast<TOut*>(s_in + PH * IPITCH);       // [NP][256*PX][CO]
  const int tid = threadIdx.x;
  const int co_off = WCONST ? wp.co_off : 0;
  const int tiles_w = (d.Wo + TW - 1) / TW, tiles_h = (d.Ho + STEM_TH - 1) / STEM_TH;
  const int tw = blockIdx.x % tiles_w, th = (blockIdx.x / tiles_w) % tiles_h, n = blockIdx.x / (tiles_w * tiles_h);
  const int oh0 = th * STEM_TH, ow0 = tw * TW;
  if (!WCONST)
    for (int i = tid; i < TAPS * CO; i += 256) s_w[i] = __ldg(d.w_f32 + i);
  for (int i = tid; i < CO; i += 256) {
    s_sc[i] = d.scale ? __ldg(d.scale + co_off + i) : 1.f;
    s_sh[i] = d.shift ? __ldg(d.shift + co_off + i) : 0.f;
  }
  for (int i = tid; i < PH * IW; i += 256) {
    const int row = i / IW, rem = i - row * IW;
    const int ih = oh0 * STRIDE - d.pad + row;
    int col, c;
    if (LAYOUT == IN_NCHW_F32) { c = rem / PW; col = rem - c * PW; }                            // column fastest: coalesced per channel plane
    else { col = rem / 3; c = rem - col * 3; }                                                  // byte order of the camera frame
    const int iw = ow0 * STRIDE - d.pad + col;
    float v = 0.f;
    if ((unsigned)ih < (unsigned)d.H && (unsigned)iw < (unsigned)d.W) {
      if (LAYOUT == IN_NCHW_F32) v = __ldg(static_cast<const float*>(d.in) + ((size_t)(n * 3 + c) * d.H + ih) * d.W + iw);
      else v = (float)__ldg(static_cast<const unsigned char*>(d.in) + ((size_t)(n * d.H + ih) * d.W + iw) * 3 + c) / 255.f;
    }
    s_in[row * IPITCH + col * 3 + c] = v;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  unsigned long long acc[PX][CO / 2];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int j = 0; j < CO / 2; ++j) acc[p][j] = 0ull;
  // filter rows: unrolled for the 3x3 stem; a real loop for 7x7 (147 taps x 16 FFMA2 would be 60 KB of straight-line code) - the
  // row index is warp-uniform, so the constant-bank weights are still fetched through uniform registers
#pragma unroll(KS <= 3 ? KS : 1)
  for (int r = 0; r < KS; ++r) {
    const float* row = s_in + (ty * STRIDE + r) * IPITCH + tx * STRIDE * 3;
#pragma unroll
    for (int qc = 0; qc < KS * 3; ++qc) {                           // (q, c) in the order of the weight rows: k = (r*KS + q)*3 + c
      unsigned long long x[PX];
#pragma unroll
      for (int p = 0; p < PX; ++p) x[p] = dup2(row[p * 16 * STRIDE * 3 + qc]);
      const ulonglong2* w2 = reinterpret_cast<const ulonglong2*>(s_w + (r * KS * 3 + qc) * CO);
      const unsigned long long* wc = reinterpret_cast<const unsigned long long*>(wp.w) + (r * KS * 3 + qc) * (CO / 2);
#pragma unroll
      for (int j = 0; j < CO / 4; ++j) {
        ulonglong2 w;
        if constexpr (WCONST) { w.x = wc[2 * j]; w.y = wc[2 * j + 1]; }
        else w = w2[j];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          acc[p][2 * j] = ffma2(x[p], w.x, acc[p][2 * j]);
          acc[p][2 * j + 1] = ffma2(x[p], w.y, acc[p][2 * j + 1]);
        }
      }
    }
  }
  constexpr int BLK_PIX = STEM_TH * TW;
  constexpr int SW_E = 16 / (int)sizeof(TOut);
  constexpr int SW_VPP = CO / SW_E;
  constexpr int SW_RPL = (128 / (CO * (int)sizeof(TOut))) > 0 ? (128 / (CO * (int)sizeof(TOut))) : 1;
  int sat = 0;
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int pl = ty * TW + tx + 16 * p;
    const int sw_x = (pl / SW_RPL) & (SW_VPP - 1);
#pragma unroll
    for (int j4 = 0; j4 < CO; j4 += 4) {
      const float2 a = *reinterpret_cast<const float2*>(&acc[p][j4 / 2]), b = *reinterpret_cast<const float2*>(&acc[p][j4 / 2 + 1]);
      const float raw[4] = {a.x, a.y, b.x, b.y};
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float y = fmaf(raw[j], s_sc[j4 + j], s_sh[j4 + j]);
        if (d.act == ACT_LEAKY) y = y > 0.f ? y : 0.1f * y;
        else if (d.act == ACT_RELU) y = fmaxf(y, 0.f);
        v[j] = y;
      }
      const int vi = j4 / SW_E, within = j4 - vi * SW_E;
      store4f<FO>(s_out + (size_t)pl * CO + ((vi ^ sw_x) * SW_E) + within, (long long)BLK_PIX * CO, v, sat);
    }
  }
  // (pixels of a ragged tile outside the image are computed on zero padding and never stored; they cannot overflow the format either)
  if (sat && d.sat_flag) atomicOr(d.sat_flag, YOLO_SAT_ACT_FFMA);
  __syncthreads();
#pragma unroll
  for (int q = 0; q < FO::NP; ++q) {
    const uint4* src = reinterpret_cast<const uint4*>(s_out + (size_t)q * BLK_PIX * CO);
    TOut* dst = static_cast<TOut*>(d.out) + (size_t)q * d.out_plane_stride + d.out_coff + co_off;
    for (int i = tid; i < BLK_PIX * SW_VPP; i += 256) {
      const int pix = i / SW_VPP, vi = i - pix * SW_VPP;
      const int oh = oh0 + pix / TW, ow = ow0 + pix % TW;
      if (oh < d.Ho && ow < d.Wo)
        *reinterpret_cast<uint4*>(dst + ((size_t)(n * d.Ho + oh) * d.Wo + ow) * d.out_cpitch + vi * SW_E) =
            src[pix * SW_VPP + (vi ^ ((pix / SW_RPL) & (SW_VPP - 1)))];
    }
  }
}

template <class FO, int CO, int PX, int KS, int STRIDE>
static int launch_stem_tile_t(const ConvKArgs& a, int in_layout, cudaStream_t st) {
  const ConvDesc& d = a.d;
  constexpr int TW = 16 * PX, TAPS = KS * KS * 3;
  constexpr int PH = (STEM_TH - 1) * STRIDE + KS, PW = (TW - 1) * STRIDE + KS, IW = PW * 3, IPITCH = ((IW + 15) / 32) * 32 + 16;
  const int out_bytes = FO::NP * STEM_TH * TW * CO * (int)sizeof(typename FO::T);
  const int blocks = d.N * ((d.Ho + STEM_TH - 1) / STEM_TH) * ((d.Wo + TW - 1) / TW);
  static const bool const_off = [] { const char* e = getenv("YOLO_B200_STEM_WCONST"); return e && e[0] == '0'; }();
  if (d.w_host && !const_off) {
    const int smem = (2 * CO + PH * IPITCH) * 4 + out_bytes;
    int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&stem_tile_kernel<FO, IN_NCHW_F32, CO, PX, true, KS, STRIDE>), 160 * 1024);
    if (!rc) rc = ensure_dyn_smem(reinterpret_cast<const void*>(&stem_tile_kernel<FO, IN_NHWC_U8, CO, PX, true, KS, STRIDE>), 160 * 1024);
    if (rc) return rc;
    StemWeights<TAPS * CO> wp;
    for (int co = 0; co < d.Cout; co += CO) {                     // one launch per CO output channels (the 64-channel 7x7 stem: two)
      for (int k = 0; k < TAPS; ++k) memcpy(wp.w + k * CO, d.w_host + (size_t)k * d.cout_pad + co, sizeof(float) * CO);
      wp.co_off = co;
      if (in_layout == IN_NCHW_F32) stem_tile_kernel<FO, IN_NCHW_F32, CO, PX, true, KS, STRIDE><<<blocks, 256, smem, st>>>(a, wp);
      else stem_tile_kernel<FO, IN_NHWC_U8, CO, PX, true, KS, STRIDE><<<blocks, 256, smem, st>>>(a, wp);
      if (co) ++g_launches;
    }
    return YOLO_OK;
  }
  if (d.Cout != CO) return fail(YOLO_E_UNSUPPORTED, "stem: the shared-memory-weights variant takes Cout == %d", CO);
  const int smem = (TAPS * CO + 2 * CO + PH * IPITCH) * 4 + out_bytes;
  StemWeights<1> none;
  none.co_off = 0;
  int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&stem_tile_kernel<FO, IN_NCHW_F32, CO, PX, false, KS, STRIDE>), 160 * 1024);
  if (!rc) rc = ensure_dyn_smem(reinterpret_cast<const void*>(&stem_tile_kernel<FO, IN_NHWC_U8, CO, PX, false, KS, STRIDE>), 160 * 1024);
  if (rc) return rc;
  if (in_layout == IN_NCHW_F32) stem_tile_kernel<FO, IN_NCHW_F32, CO, PX, false, KS, STRIDE><<<blocks, 256, smem, st>>>(a, none);
  else stem_tile_kernel<FO, IN_NHWC_U8, CO, PX, false, KS, STRIDE><<<blocks, 256, smem, st>>>(a, none);
  return YOLO_OK;
}
template <class FO>
static int launch_stem_tile(const ConvKArgs& a, int in_layout, cudaStream_t st) {
  // one pixel per thread: 64 registers and 40 KB per block -> four blocks per SM hide the load and store phases of each other
  // (two pixels per thread reuse every weight load twice but fit one or two blocks: YOLO_B200_STEM_PX=2, measured slower)
  if (a.d.kh == 7) return launch_stem_tile_t<FO, 32, 1, 7, 2>(a, in_layout, st);      // DenseNet stem: 7x7 / 2, 64 channels in two halves
  static const bool px2 = [] { const char* e = getenv("YOLO_B200_STEM_PX"); return e && e[0] == '2'; }();
  if (px2) return a.d.Cout == 32 ? launch_stem_tile_t<FO, 32, 2, 3, 1>(a, in_layout, st) : launch_stem_tile_t<FO, 16, 2, 3, 1>(a, in_layout, st);
  return a.d.Cout == 32 ? launch_stem_tile_t<FO, 32, 1, 3, 1>(a, in_layout, st) : launch_stem_tile_t<FO, 16, 1, 3, 1>(a, in_layout, st);
}

template <class FO, int PX>
static int launch_stem_t(const ConvKArgs& a, int in_layout, int smem, cudaStream_t st) {
  const int blocks = (a.M + 256 * PX - 1) / (256 * PX);
  // staged 16-bit planes need more than the default 48 KB
  int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&stem3x3_kernel<FO, IN_NCHW_F32, PX>), 160 * 1024);
  if (!rc) rc = ensure_dyn_smem(reinterpret_cast<const void*>(&stem3x3_kernel<FO, IN_NHWC_U8, PX>), 160 * 1024);
  if (rc) return rc;
  if (in_layout == IN_NCHW_F32) stem3x3_kernel<FO, IN_NCHW_F32, PX><<<blocks, 256, smem, st>>>(a);
  else stem3x3_kernel<FO, IN_NHWC_U8, PX><<<blocks, 256, smem, st>>>(a);
  return YOLO_OK;
}

static bool stem7_eligible(const ConvDesc& d) {           // DenseNet stem (7x7 / 2 / pad 3, 64 channels): tiled kernel with kernel-parameter weights only
  return d.kh == 7 && d.kw == 7 && d.stride == 2 && d.pad == 3 && d.Cout % 32 == 0 && d.Cout <= 64 && d.w_host != nullptr &&
         d.out_dtype != DT_F32 && d.out_dtype != DT_BF16X3;
}
bool stem_eligible(const ConvDesc& d, int in_layout) {
  if (!((in_layout == IN_NCHW_F32 || in_layout == IN_NHWC_U8) && d.Cin == 3 && !d.res && !d.pre_scale && !d.upsample2 && !d.out_nchw &&
        ((d.out_cpitch | d.out_coff) & 7) == 0 && d.cout_pad == d.Cout))
    return false;
  if (stem7_eligible(d)) return true;
  return d.kh == 3 && d.kw == 3 && d.Cout % STEM_CO == 0 && d.Cout <= 32;
}

int launch_stem(const ConvDesc& d, int in_layout, cudaStream_t st) {
  ConvKArgs a;
  a.d = d;
  a.in_layout = in_layout;
  a.M = d.N * d.Ho * d.Wo;
  a.K = d.kh * d.kw * 3;
  const int esz = d.out_dtype == DT_F32 ? 4 : 2, npl = dtype_planes(d.out_dtype);
  auto smem_for = [&](int px) { return (27 * d.cout_pad + 2 * d.Cout) * 4 + npl * 256 * px * d.Cout * esz; };   // weights, scale/shift, staged rows
  const bool two = smem_for(2) <= 150 * 1024;                 // two pixels per thread whenever the staged rows fit
  int rc;
  static const bool tile_off = [] { const char* e = getenv("YOLO_B200_STEM_TILE"); return e && e[0] == '0'; }();
  const size_t osz = (size_t)esz;
  const bool k7 = stem7_eligible(d);
  const bool tiled = ((!tile_off && d.kh == 3 && d.stride == 1 && d.pad == 1 && d.Ho == d.H && d.Wo == d.W && (d.Cout == 16 || d.Cout == 32)) || k7) &&
                     (d.Cout * osz) % 16 == 0 && (d.out_cpitch * osz) % 16 == 0 && (d.out_coff * osz) % 16 == 0 &&
                     ((size_t)d.out_plane_stride * osz) % 16 == 0 && (reinterpret_cast<uintptr_t>(d.out) & 15) == 0;
  if (tiled) {
    switch (d.out_dtype) {
      case DT_F32: rc = launch_stem_tile<FmtF32>(a, in_layout, st); break;
      case DT_BF16: rc = launch_stem_tile<FmtBF16>(a, in_layout, st); break;
      case DT_BF16X3: rc = launch_stem_tile<FmtBF16X3>(a, in_layout, st); break;
      default: rc = launch_stem_tile<FmtF16X2>(a, in_layout, st); break;
    }
    if (rc) return rc;
    ++g_launches;
    YB_CUDA(cudaGetLastError());
    return YOLO_OK;
  }
  if (k7) return fail(YOLO_E_UNSUPPORTED, "stem: 7x7 stem output slices must be 16-byte aligned");
  switch (d.out_dtype) {
    case DT_F32: rc = two ? launch_stem_t<FmtF32, 2>(a, in_layout, smem_for(2), st) : launch_stem_t<FmtF32, 1>(a, in_layout, smem_for(1), st); break;
    case DT_BF16: rc = two ? launch_stem_t<FmtBF16, 2>(a, in_layout, smem_for(2), st) : launch_stem_t<FmtBF16, 1>(a, in_layout, smem_for(1), st); break;
    case DT_BF16X3: rc = two ? launch_stem_t<FmtBF16X3, 2>(a, in_layout, smem_for(2), st) : launch_stem_t<FmtBF16X3, 1>(a, in_layout, smem_for(1), st); break;
    default: rc = two ? launch_stem_t<FmtF16X2, 2>(a, in_layout, smem_for(2), st) : launch_stem_t<FmtF16X2, 1>(a, in_layout, smem_for(1), st); break;
  }
  if (rc) return rc;
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

// ---- pooling (NHWC, channel-contiguous threads) -------------------------------------------------------
template <class F, bool IS_MAX>
__global__ void pool_kernel(const typename F::T* __restrict__ in, typename F::T* __restrict__ out, int N, int H, int W, int C, int in_cpitch,
                            int in_coff, long long in_ps, int out_cpitch, int out_coff, long long out_ps, int Ho, int Wo, int k,
                            int stride, int pad) {
  size_t total = (size_t)N * Ho * Wo * C;
  int sat = 0;                                   // pooled values never exceed their (already representable) inputs
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    size_t pix = idx / C;
    int ow = (int)(pix % Wo);
    size_t t = pix / Wo;
    int oh = (int)(t % Ho), n = (int)(t / Ho);
    float acc = IS_MAX ? -CUDART_INF_F : 0.f;
    for (int r = 0; r < k; ++r) {
      int ih = oh * stride - pad + r;
      if ((unsigned)ih >= (unsigned)H) continue;
      for (int s = 0; s < k; ++s) {
        int iw = ow * stride - pad + s;
        if ((unsigned)iw >= (unsigned)W) continue;
        float v = ld1f<F>(in + ((size_t)(n * H + ih) * W + iw) * in_cpitch + in_coff + c, in_ps);
        acc = IS_MAX ? fmaxf(acc, v) : acc + v;
      }
    }
    if (!IS_MAX) acc = acc / (float)(k * k);     // gluon AvgPool2D: count_include_pad, no padding used here
    store1f<F>(out + pix * out_cpitch + out_coff + c, out_ps, acc, sat);
  }
}

// eight channels per thread (16-byte loads per plane) when the channel slices allow it: the scalar kernel above issues one 2-byte
// load per plane, tap and element (DenseNet stem max-pool at 160 x 256 x 64, batch 64: 1.02 ms -> the vector kernel)
template <class F, bool IS_MAX>
__global__ void __launch_bounds__(256)
pool_vec8_kernel(const typename F::T* __restrict__ in, typename F::T* __restrict__ out, int N, int H, int W, int C, int in_cpitch, int in_coff,
                 long long in_ps, int out_cpitch, int out_coff, long long out_ps, int Ho, int Wo, int k, int stride, int pad) {
  const int c8 = C >> 3;
  const size_t total = (size_t)N * Ho * Wo * c8;
  int sat = 0;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c8) * 8;
    const size_t pix = idx / c8;
    const int ow = (int)(pix % Wo);
    const size_t t = pix / Wo;
    const int oh = (int)(t % Ho), n = (int)(t / Ho);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = IS_MAX ? -CUDART_INF_F : 0.f;
    for (int r = 0; r < k; ++r) {
      const int ih = oh * stride - pad + r;
      if ((unsigned)ih >= (unsigned)H) continue;
      for (int s = 0; s < k; ++s) {
        const int iw = ow * stride - pad + s;
        if ((unsigned)iw >= (unsigned)W) continue;
        float v[8];
        load8f<F>(in + ((size_t)(n * H + ih) * W + iw) * in_cpitch + in_coff + c, in_ps, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = IS_MAX ? fmaxf(acc[j], v[j]) : acc[j] + v[j];
      }
    }
    if (!IS_MAX) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = acc[j] / (float)(k * k);
    }
    typename F::T* op = out + pix * out_cpitch + out_coff + c;
    store4f<F>(op, out_ps, acc, sat);
    store4f<F>(op + 4, out_ps, acc + 4, sat);
  }
}

int launch_pool(const void* in, void* out, int dtype, int N, int H, int W, int C, int in_cpitch, int in_coff,
                long long in_plane_stride, int out_cpitch, int out_coff, long long out_plane_stride, int k, int stride, int pad,
                int is_max, cudaStream_t st) {
  int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  size_t total = (size_t)N * Ho * Wo * C;
  if (total == 0) return fail(YOLO_E_SHAPE, "pool: empty problem");
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const bool vec8 = dtype != DT_F32 && C % 8 == 0 && ((in_cpitch | in_coff | out_cpitch | out_coff) & 7) == 0 && in_plane_stride % 8 == 0 &&
                    out_plane_stride % 8 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec8) {
    blocks = (int)((total / 8 + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
#define YB_POOLV(F, MX)                                                                                                         \
  pool_vec8_kernel<F, MX><<<blocks, 256, 0, st>>>(static_cast<const F::T*>(in), static_cast<F::T*>(out), N, H, W, C, in_cpitch, \
                                                  in_coff, in_plane_stride, out_cpitch, out_coff, out_plane_stride, Ho, Wo, k, stride, pad)
    if (dtype == DT_BF16) { if (is_max) YB_POOLV(FmtBF16, true); else YB_POOLV(FmtBF16, false); }
    else if (dtype == DT_BF16X3) { if (is_max) YB_POOLV(FmtBF16X3, true); else YB_POOLV(FmtBF16X3, false); }
    else { if (is_max) YB_POOLV(FmtF16X2, true); else YB_POOLV(FmtF16X2, false); }
#undef YB_POOLV
    ++g_launches;
    YB_CUDA(cudaGetLastError());
    return YOLO_OK;
  }
#define YB_POOL(F, MX)                                                                                                   \
  pool_kernel<F, MX><<<blocks, 256, 0, st>>>(static_cast<const F::T*>(in), static_cast<F::T*>(out), N, H, W, C, in_cpitch, \
                                             in_coff, in_plane_stride, out_cpitch, out_coff, out_plane_stride, Ho, Wo, k, stride, pad)
  if (dtype == DT_F32) { if (is_max) YB_POOL(FmtF32, true); else YB_POOL(FmtF32, false); }
  else if (dtype == DT_BF16) { if (is_max) YB_POOL(FmtBF16, true); else YB_POOL(FmtBF16, false); }
  else if (dtype == DT_BF16X3) { if (is_max) YB_POOL(FmtBF16X3, true); else YB_POOL(FmtBF16X3, false); }
  else { if (is_max) YB_POOL(FmtF16X2, true); else YB_POOL(FmtF16X2, false); }
#undef YB_POOL
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

// ---- DenseNet pre-activation (BN -> ReLU) as its own pass ---------------------------------------------------------------------------
// In a dense block every layer applies ITS OWN BatchNorm to the shared concatenated features, so the activation cannot be folded into
// the producer's epilogue.  The FFMA kernel applies it while gathering; the tensor-core kernel stages raw tiles by TMA, so the
// pre-activated input is materialised once per layer (channel-padded to the K step with zeros) and the convolution runs unmodified.
template <class F, bool ALIGNED>
__global__ void __launch_bounds__(256)
preact_kernel(const typename F::T* __restrict__ in, typename F::T* __restrict__ out, long long pixels, int C, int Cpad, int in_cpitch, int in_coff,
              long long in_ps, int out_cpitch, long long out_ps, const float* __restrict__ scale, const float* __restrict__ shift, int* sat_flag) {
  const int oct = Cpad >> 3;
  const long long total = pixels * oct;
  int sat = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / oct;
    const int c = (int)(i - m * oct) * 8;
    float v[8];
    if (ALIGNED) {
      if (c < C) {
        load8f<F>(in + m * in_cpitch + in_coff + c, in_ps, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], __ldg(scale + c + j), __ldg(shift + c + j)), 0.f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
    } else {
      // concat buffers whose channel count is not a multiple of 8 (DenseNet block 4 starts at 260 channels): element loads; the copy
      // this pass writes is aligned, which is what puts the convolution behind it on the tensor cores
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cj = c + j;
        v[j] = cj < C ? fmaxf(fmaf(ld1f<F>(in + m * in_cpitch + in_coff + cj, in_ps), __ldg(scale + cj), __ldg(shift + cj)), 0.f) : 0.f;
      }
    }
    store4f<F>(out + m * out_cpitch + c, out_ps, v, sat);
    store4f<F>(out + m * out_cpitch + c + 4, out_ps, v + 4, sat);
  }
  if (sat && sat_flag) atomicOr(sat_flag, YOLO_SAT_ACT_FFMA);
}

int launch_preact(const void* in, void* out, int dtype, long long pixels, int C, int Cpad, int in_cpitch, int in_coff, long long in_plane_stride,
                  int out_cpitch, long long out_plane_stride, const float* scale, const float* shift, int* sat_flag, cudaStream_t st) {
  if (Cpad % 8 || out_cpitch % 8 || out_plane_stride % 8 || dtype == DT_F32) return fail(YOLO_E_UNSUPPORTED, "preact: needs 16-bit planes and a padded channel count % 8 == 0");
  const bool aligned = C % 8 == 0 && in_cpitch % 8 == 0 && in_coff % 8 == 0 && in_plane_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  const long long total = pixels * (Cpad / 8);
  if (total <= 0) return fail(YOLO_E_SHAPE, "preact: empty problem");
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
#define YB_PREACT(F)                                                                                                                             \
  do {                                                                                                                                           \
    if (aligned) preact_kernel<F, true><<<blocks, 256, 0, st>>>(static_cast<const F::T*>(in), static_cast<F::T*>(out), pixels, C, Cpad, in_cpitch, \
                                                                in_coff, in_plane_stride, out_cpitch, out_plane_stride, scale, shift, sat_flag);  \
    else preact_kernel<F, false><<<blocks, 256, 0, st>>>(static_cast<const F::T*>(in), static_cast<F::T*>(out), pixels, C, Cpad, in_cpitch,       \
                                                         in_coff, in_plane_stride, out_cpitch, out_plane_stride, scale, shift, sat_flag);         \
  } while (0)
  if (dtype == DT_BF16) YB_PREACT(FmtBF16);
  else if (dtype == DT_BF16X3) YB_PREACT(FmtBF16X3);
  else YB_PREACT(FmtF16X2);
#undef YB_PREACT
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

}  // namespace yb
