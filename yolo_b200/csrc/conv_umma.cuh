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
  void* w_packed = nullptr;   // device: [planes][rows][kdim] 16-bit, K-major, k = (r*kw+s)*Cin + c
  size_t w_bytes = 0;
  int rows = 0;               // weight rows per plane (Cout padded to the N tile)
  size_t kdim = 0;            // packed K (taps * Cin; taps * 64 for the interleaved Cin = 32 layout)
  alignas(64) unsigned char map_a[128];   // CUtensorMap (im2col) for the activations
  alignas(64) unsigned char map_b[128];   // CUtensorMap (tiled)  for the weights
  alignas(64) unsigned char map_bw[128];  // same with a 256-row box (128 x 256 tiles)
  bool has_map_bw = false;
  bool a_tiled = false;       // 1x1 conv: map_a is a tiled 2-D map over [pixels][channels]
  long long a_plane_rows = 0;
  int max_batch = 0;
  bool pad_high_full = false; // im2col: pad = 0 on the near edge, k-1 on the far edge (parity classes of a strided data gradient)
  bool c32i = false;          // Cin = 32, plane-interleaved activations ([hi(32) | lo(32)] per pixel): KIND 7
  float prescale = 1.f;      // 2^s applied to the weights when they are packed (fp16 planes stay normal)
  float acc_scale = 1.f;     // 2^-s, undone in the epilogue
  mutable void* split_ws = nullptr;   // split-K partial tiles of this layer's tail wave (grown on first use)
  mutable size_t split_ws_bytes = 0;
};

// optional epilogue extras of the training step
struct UmmaExtra {
  const float* acc_scale_dev = nullptr;   // second accumulator scale read from device memory (dynamic gradient scale)
  int accum = 0;                          // fp32 output: out += result
  float* stats = nullptr;                 // [umma_stats_groups(M)][2][Cout] per-warp column sums / sums of squares of the result
  size_t stats_groups_out = 0;            // set by the launcher: row groups (of 32 pixels) the kernel wrote statistics for
};

// w_oihw may be null: the packed planes are then zero-filled and written later by umma_pack_device
int umma_prepare_weights(UmmaConv& u, int precision, const float* w_oihw, int cout, int cin, int kh, int kw, int stride,
                         int pad, int in_dtype, bool has_prologue, int out_nchw, bool in_interleaved, cudaStream_t st,
                         bool rect_ok = false);
// power-of-two weight pre-scale so that wmax lands in [2^(top-1), 2^top) (fp16x3 only)
void umma_set_prescale(UmmaConv& u, float wmax, int top);
bool umma_disabled();                                             // YOLO_B200_DISABLE_UMMA
// (re)pack the planes on the device from the fp32 [kh*kw*w_cin][cout_pad] matrix (k = tap*w_cin + c) of a w_cin -> w_cout convolution, in
// one launch: `fwd` is that convolution (may be null / not eligible), `dg` its data-gradient convolution (n_dg = 1: w_cout -> w_cin with
// a flipped filter) or the four parity classes py*2 + px of a 3x3 / stride 2 / pad 1 layer (n_dg = 4: dg[q].kh = 1 + py, .kw = 1 + px).
// *sat_flag |= 2 when a scaled weight leaves the fp16 range.
int umma_pack_device(const UmmaConv* fwd, const UmmaConv* dg, int n_dg, const float* w_mat, int w_cin, int w_cout, int cout_pad, int kh, int kw,
                     int* sat_flag, cudaStream_t st);
int umma_build_maps(UmmaConv& u, void* in_base, int max_batch, int H, int W, int C, int cpitch, int coff);
int launch_conv_umma(const UmmaConv& u, const ConvDesc& d, cudaStream_t st, UmmaExtra* ex = nullptr);
size_t umma_stats_groups(int M);
void umma_release(UmmaConv& u);

int device_sm_count(int* sms);                                    // per-device cache
int load_tma_entry_points(void** encode_tiled, void** encode_im2col);

}  // namespace yb
