// tcgen05 weight-gradient convolution (training step): host-side state and entry points.
#pragma once
#include "common.cuh"

namespace yb {

struct WgradPlan {
  bool enabled = false;
  int cin = 0, cout = 0, kh = 1, kw = 1, stride = 1, pad = 0;
  int H = 0, W = 0, Ho = 0, Wo = 0;
  int in_coff = 0;
  int x_plane_n = 0;              // images per plane of the activation buffer (forward workspace: max_batch)
  long long dz_plane_rows = 0;    // pixel rows per plane of the dz matrix
  int bn = 0;                     // N tile (<= 256, multiple of 64)
  bool il32 = false;              // Cin = 32 with plane-interleaved activations ([hi(32) | lo(32)] per pixel): the transposed kernel
  alignas(64) unsigned char map_x[128];    // im2col map over the layer input, box = 64 pixels x 64 channels
  alignas(64) unsigned char map_dz[128];   // tiled map over dz [planes * dz_plane_rows][Cout], box = 64 pixels x 64 channels
};

bool wgrad_umma_eligible(int cin, int cout, int kh, int kw, int in_dtype, bool in_interleaved);
// Cin = 32, plane-interleaved input (cpitch 64): dW^T = dz^T x through the transposed kernel (Cout = 64 or a multiple of 128)
bool wgrad_umma_il32_eligible(int cin, int cout, int kh, int kw, int in_dtype, bool in_interleaved, int cpitch, int coff);
// x_base: first plane of the layer input buffer (NHWC, cpitch channels per pixel, planes x_plane_n images apart);
// dz_base: [2][dz_plane_rows][cout] fp16 planes of the (scaled) pre-activation gradient, rows >= M zero up to the next multiple of 64.
int wgrad_umma_plan(WgradPlan& w, void* x_base, int x_plane_n, int H, int W, int cin, int cpitch, int coff, int kh, int kw, int stride, int pad,
                    void* dz_base, long long dz_plane_rows, int cout);
// scratch bytes needed for `batch` images (split partial sums; 0 when the layer runs unsplit)
size_t wgrad_umma_scratch_bytes(const WgradPlan& w, int batch, int cout_pad, int num_sms);
// dW[k][n] (row pitch cout_pad, k = tap*cin + c) = acc_scale * (*acc_scale_dev) * sum_m x[m][k] * dz[m][n]
int launch_wgrad_umma(const WgradPlan& w, int batch, float* dW, int cout_pad, float acc_scale, const float* acc_scale_dev, float* scratch,
                      size_t scratch_bytes, int variant, cudaStream_t st);

// fp32 -> two fp16 planes (v*scale = hi + lo); dst rows have `dst_pitch` elements, planes `plane_stride` elements apart
int launch_split_f16x2(const float* src, long long rows, int cols, int src_pitch, float scale, const float* scale_dev, void* dst, int dst_pitch,
                       long long plane_stride, int* sat_flag, cudaStream_t st);

}  // namespace yb
