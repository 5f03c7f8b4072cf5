// Shared helpers for the yolo_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/yolo_b200.h"

namespace yb {

// Thread-local text of the last failure for calls that have no handle (decode entry points, create()).
std::string& tls_error();
int fail(int code, const char* fmt, ...);

#define YB_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return yb::fail(YOLO_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,            \
                      cudaGetErrorString(_e));                                            \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device, per-function attribute: set once per (function, device) - the C ABI
// allows handles on several devices in one process (reference: one executor per context, yolo_gluon.get_ctx).
int ensure_dyn_smem(const void* func, int bytes);

// Global launch counter (kernels of THIS library), read by yolo_last_launch_count().
extern thread_local int g_launches;

enum Act : int { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2 };
// DT_BF16X3: three bf16 planes v = v0 + v1 + v2 (exact split of an fp32 value), plane p at element offset p*plane_stride
// DT_F16X2:  two fp16 planes  v = v0 + v1 (22-bit split).  v1 is stored UNSCALED so that all plane products of the tensor-core
//            path can accumulate in one accumulator; for activations of O(1) its fp16 subnormal spacing (6e-8) is below fp32
//            epsilon, and the weights are pre-scaled by a power of two (conv_umma.cu) so that their low plane stays normal.
enum DType : int { DT_F32 = 0, DT_BF16 = 1, DT_BF16X3 = 2, DT_F16X2 = 3 };
constexpr float kF16LoScale = 1.f, kF16LoScaleInv = 1.f, kF16Max = 65504.f;   // low plane stored unscaled (see above)
inline int dtype_planes(int dt) { return dt == DT_BF16X3 ? 3 : (dt == DT_F16X2 ? 2 : 1); }
inline int dtype_bytes_per_elem(int dt) { return dt == DT_F32 ? 4 : 2 * dtype_planes(dt); }

// One convolution (+ fused prologue / epilogue) as executed by the kernels.  Activations are NHWC;
// a tensor may be a channel slice [coff, coff+C) of a wider buffer with `cpitch` channels per pixel
// (this is how route concat and DenseNet concat are realised without copies).
struct ConvDesc {
  // input
  const void* in;  int in_dtype;  int N, H, W, Cin;  int in_cpitch, in_coff;  long long in_plane_stride;
  // filter
  int kh, kw, stride, pad;  int Cout;
  int in_dil;                  // >1: the input is implicitly zero-dilated (data-gradient of a strided conv); 0/1 = off
  const float* w_f32;          // [kh*kw*Cin][CoutPad4] fp32 (SIMT path), k = (r*kw+s)*Cin + c
  int cout_pad;                // row pitch of w_f32
  const float* w_host;         // host copy of w_f32 when it is static (inference stem: weights travel as kernel parameters), else null
  // prologue on the input (DenseNet pre-activation BN->ReLU): v = relu(v*pre_scale[c] + pre_shift[c])
  const float* pre_scale;  const float* pre_shift;
  // epilogue: y = act(acc*scale[o] + shift[o]) (+ residual)
  const float* scale;  const float* shift;  int act;
  const void* res;  int res_cpitch, res_coff;  long long res_plane_stride;   // same dtype as out
  // output
  void* out;  int out_dtype;  int Ho, Wo;  int out_cpitch, out_coff;  long long out_plane_stride;
  int upsample2;               // write every output pixel to the 2x2 block of a (2Ho, 2Wo) map
  int out_nchw;                // fp32 only: store as (N, Cout, Ho, Wo)
  const float* dyn_scale;      // optional device scalar multiplied into the accumulator before scale/shift (dynamic gradient scale)
  int* sat_flag;               // optional device flag: |= 1 when a value leaves the fp16 range of the DT_F16X2 high plane
};

// in_layout: 0 = NHWC of in_dtype, 1 = NCHW fp32 (stem only), 2 = NHWC uint8 scaled by 1/255 (stem only)
int launch_conv_simt(const ConvDesc& d, int in_layout, cudaStream_t st);
bool stem_eligible(const ConvDesc& d, int in_layout);          // dedicated 3x3 kernel for the 3-channel network input
int launch_stem(const ConvDesc& d, int in_layout, cudaStream_t st);
// y[c] = relu(x[c]*scale[c] + shift[c]) for c < C, 0 for C <= c < Cpad (DenseNet pre-activation BN -> ReLU, materialised once per layer so
// that the convolution behind it is a plain tensor-core convolution); 16-bit plane formats, C % 8 == 0
int launch_preact(const void* in, void* out, int dtype, long long pixels, int C, int Cpad, int in_cpitch, int in_coff, long long in_plane_stride,
                  int out_cpitch, long long out_plane_stride, const float* scale, const float* shift, int* sat_flag, cudaStream_t st);
int launch_pool(const void* in, void* out, int dtype, int N, int H, int W, int C, int in_cpitch, int in_coff,
                long long in_plane_stride, int out_cpitch, int out_coff, long long out_plane_stride, int k, int stride, int pad,
                int is_max, cudaStream_t st);

}  // namespace yb
