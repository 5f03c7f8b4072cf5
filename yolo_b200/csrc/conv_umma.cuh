// tcgen05 (UMMA) implicit-GEMM convolution: host-side state and entry points.
#pragma once
#include "common.cuh"

namespace yb {

struct UmmaConv {
  bool eligible = false;      // shape/precision can run on the tensor-core path
  bool enabled = false;       // tensor maps built for the current workspace
  int precision = 0;
  int cout = 0, cin = 0, kh = 1, kw = 1, stride = 1, pad = 0;
  int bn_tile = 0;            // N tile (UMMA N)
  int bk = 64;                // K elements per pipeline stage: 64 (128-byte swizzle rows) or 32 (64-byte rows, Cin % 64 != 0)
  void* w_packed = nullptr;   // device: [planes][cout_padded][K] 16-bit, K-major, k = (r*kw+s)*Cin + c
  size_t w_bytes = 0;
  alignas(64) unsigned char map_a[128];   // CUtensorMap (im2col) for the activations
  alignas(64) unsigned char map_b[128];   // CUtensorMap (tiled)  for the weights
  alignas(64) unsigned char map_b2[128];  // same with a half-height box (2-CTA path)
  bool has_map_b2 = false;
  alignas(64) unsigned char map_bw[128];  // same with a 256-row box (128 x 256 tiles)
  bool has_map_bw = false;
  bool a_tiled = false;       // 1x1 conv: map_a is a tiled 2-D map over [pixels][channels]
  long long a_plane_rows = 0;
  int max_batch = 0;
  bool c32i = false;          // Cin = 32, plane-interleaved activations ([hi(32) | lo(32)] per pixel): KIND 7
  float acc_scale = 1.f;     // 2^-s: the packed weights are pre-scaled by 2^s (fp16 planes stay normal), undone in the epilogue
};

int umma_prepare_weights(UmmaConv& u, int precision, const float* w_oihw, int cout, int cin, int kh, int kw, int stride,
                         int pad, int in_dtype, bool has_prologue, int out_nchw, bool in_interleaved, cudaStream_t st);
int umma_build_maps(UmmaConv& u, void* in_base, int max_batch, int H, int W, int C, int cpitch, int coff);
int launch_conv_umma(const UmmaConv& u, const ConvDesc& d, cudaStream_t st);
void umma_release(UmmaConv& u);

}  // namespace yb
