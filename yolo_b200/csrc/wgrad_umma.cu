// Weight gradient of a convolution on the 5th-generation tensor cores (training step).
//
// Replaces cuDNN's convolution-backward-filter dispatched by MXNet autograd in `sum(losses).backward()` (reference
// car/YOLO.py:394) for the gluoncv `_conv2d` blocks (yolo_modules/basic_yolo.py:20-26,118-121).
//
//   dW[k][n] = sum_m X[m][k] * dZ[m][n],   m = output pixel (batch*Ho*Wo), k = (r*kw + s)*Cin + c, n = output channel
//
// The contraction runs over PIXELS.  Both operands live in HBM pixel-major (NHWC activations, [pixels][Cout] gradients), i.e. with
// the contraction index as the slow one: "MN-major" operands in tcgen05 terms.  They are staged exactly as they lie - the same
// im2col TMA map as the forward convolution gathers 64 pixels x 64 channels of one filter tap, a tiled map fetches 64 pixels x
// 64 channels of dz - and the MMA reads them TRANSPOSED through MN-major shared-memory descriptors (instruction-descriptor bits
// 15/16).  No transposed copy of x or dz ever exists.
//
// D tile = 128 k-rows (two (tap, 64-channel) units) x BN output channels, accumulated over the pixel blocks of one SPLIT of the
// pixel range; splits write partial tiles that a second kernel adds in a FIXED order (deterministic - no atomics).
// fp32-grade arithmetic like the forward kernel (conv_umma.cu): x = hi + lo, dz = hi + lo fp16 planes, products hi*hi, hi*lo, lo*hi;
// two-level accumulation (short TMEM partials added into fp32 registers with round-to-nearest; correction products first).
//
// Warps: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..5 / 6..9 = accumulation groups (BN/2 columns each).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <string.h>

#include "umma_ptx.cuh"
#include "wgrad_umma.cuh"
#include "conv_umma.cuh"

namespace yb {

constexpr int WG_THREADS = 320;
constexpr int WG_PIX = 64;                     // pixels per pipeline stage (MMA K = 4 x 16)
constexpr int WG_TILE_BYTES = WG_PIX * 128;    // one TMA box: 64 pixel rows x 64 channels x 2 bytes
constexpr int WG_MAX_STAGES = 6;
constexpr int WG_SMEM_LIMIT = 227 * 1024;

struct WgradParams {
  int M, n_units, cin_blocks, kw;
  int Ho, Wo, stride, pad;
  int in_coff, x_plane_n;
  long long dz_plane_rows;
  int BN, n_tiles_n, splits, stages, flush;
  int K, Cout, cout_pad;
  float acc_scale;
  const float* acc_scale_dev;
  float* out;
  long long out_split_stride;                 // elements between the partial results of two splits
  int variant;                                // descriptor-convention probe of the unit test (0 = the documented layout)
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int nb64 = p.BN >> 6;                                     // 64-channel blocks of the dz tile
  const int a_bytes = 2 * 2 * WG_TILE_BYTES;                      // planes x units
  const int stage_bytes = a_bytes + 2 * nb64 * WG_TILE_BYTES;
  const uint32_t tiles_end = smem_base + p.stages * stage_bytes;
  unsigned char* aux = smem_gen + (size_t)p.stages * stage_bytes;
  const uint32_t bar_full = tiles_end, bar_empty = tiles_end + 8 * WG_MAX_STAGES;
  const uint32_t bar_pfull = tiles_end + 16 * WG_MAX_STAGES, bar_pempty = bar_pfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 16 * WG_MAX_STAGES + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x;
  const int split = unit % p.splits, tile = unit / p.splits;
  const int kt = tile / p.n_tiles_n, nt = tile - kt * p.n_tiles_n;
  const int nblk = (p.M + WG_PIX - 1) / WG_PIX;
  const int pb_begin = (int)((long long)nblk * split / p.splits), pb_end = (int)((long long)nblk * (split + 1) / p.splits);
  const int nstage_total = pb_end - pb_begin;
  const int npart = (nstage_total + p.flush - 1) / p.flush;
  const int acc_stride = p.BN;
  const int tmem_cols = 2 * p.BN < 32 ? 32 : 2 * p.BN;            // 128 / 256 / 512

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_x);
    prefetch_tmap(&map_dz);
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_pfull + 8 * b, 1); mbar_init(bar_pempty + 8 * b, 8); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      const int HoWo = p.Ho * p.Wo;
      int tap_r[2], tap_s[2], cb[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        int U = 2 * kt + u;
        if (U >= p.n_units) U = p.n_units - 1;                    // odd unit count: the second half of the last tile is not stored
        const int tap = U / p.cin_blocks;
        cb[u] = U - tap * p.cin_blocks;
        tap_r[u] = tap / p.kw; tap_s[u] = tap - tap_r[u] * p.kw;
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        const int m0 = pb * WG_PIX;
        const int img = m0 / HoWo, rem = m0 - img * HoWo;
        const int oh = rem / p.Wo, ow = rem - oh * p.Wo;
        const int bw = ow * p.stride - p.pad, bh = oh * p.stride - p.pad;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t full = bar_full + 8 * stage;
        const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + a_bytes;
        mbar_expect_tx(full, (uint32_t)stage_bytes);
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
          for (int u = 0; u < 2; ++u)
            tma_load_im2col_4d(sa + (pl * 2 + u) * WG_TILE_BYTES, &map_x, full, p.in_coff + cb[u] * 64, bw, bh, img + pl * p.x_plane_n,
                               (uint16_t)tap_s[u], (uint16_t)tap_r[u]);
          for (int j = 0; j < nb64; ++j)
            tma_load_2d(sb + (pl * nb64 + j) * WG_TILE_BYTES, &map_dz, full, nt * p.BN + j * 64, (int)(m0 + pl * p.dz_plane_rows));
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.BN, true, 128, true, true);       // both operands MN-major
      const uint32_t lbo = p.variant == 1 ? 1024u : (uint32_t)WG_TILE_BYTES, sbo = p.variant == 1 ? (uint32_t)WG_TILE_BYTES : 1024u;
      int stage = 0;
      uint32_t phase = 0, pcount = 0;
      for (int i = 0; i < nstage_total; ++i) {
        const int pbuf = pcount & 1;
        const uint32_t tmem_main = tmem_base + pbuf * acc_stride;
        if (i % p.flush == 0) {                                             // new partial: its buffer must have been drained
          mbar_wait(bar_pempty + 8 * pbuf, ((pcount >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + a_bytes;
        uint32_t written = (i % p.flush == 0) ? 0u : 1u;
        // correction products first (lo*hi, hi*lo), the leading hi*hi product closes the stage (merged accumulation, conv_umma.cu)
        constexpr int PA[3] = {1, 0, 0}, PB[3] = {0, 1, 0};
#pragma unroll
        for (int pi = 0; pi < 3; ++pi) {
          const uint64_t adesc = make_smem_desc_mn(sa + PA[pi] * 2 * WG_TILE_BYTES, lbo, sbo);
          const uint64_t bdesc = make_smem_desc_mn(sb + PB[pi] * nb64 * WG_TILE_BYTES, lbo, sbo);
#pragma unroll
          for (int k = 0; k < WG_PIX / UMMA_K; ++k) {
            const uint64_t koff = (uint64_t)((k * UMMA_K * 128) >> 4);      // 16 pixel rows of 128 bytes
            umma_bf16(tmem_main, adesc + koff, bdesc + koff, idesc, written);
            written = 1;
          }
        }
        umma_commit(bar_empty + 8 * stage);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if ((i + 1) % p.flush == 0 || i + 1 == nstage_total) { umma_commit(bar_pfull + 8 * pbuf); ++pcount; }
      }
    }
  } else {
    // =========================== accumulation groups + epilogue ===========================
    const int group = (warp - 2) >> 2;
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;                         // D row = k index within the tile
    const int gcols = p.BN >> 1;                                  // columns of this group (32, 64 or 128)
    const int nchunks = gcols >> 5;
    const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
    float acc[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[c][i] = 0.f;
    for (int part = 0; part < npart; ++part) {
      const int pbuf = part & 1;
      mbar_wait(bar_pfull + 8 * pbuf, (part >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_addr + pbuf * acc_stride + group * gcols;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nchunks) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[c][i] += __uint_as_float(v[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pempty + 8 * pbuf);
    }
    const int U = 2 * kt + (row >> 6);
    if (U < p.n_units) {
      const float sc = p.acc_scale * (p.acc_scale_dev ? __ldg(p.acc_scale_dev) : 1.f);
      const int k = U * 64 + (row & 63);
      float* op = p.out + (size_t)split * p.out_split_stride + (size_t)k * p.cout_pad + nt * p.BN + group * gcols;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nchunks) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(op + c * 32 + i) = make_float4(acc[c][i] * sc, acc[c][i + 1] * sc, acc[c][i + 2] * sc, acc[c][i + 3] * sc);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---- Cin = 32, plane-interleaved activations: the transposed problem -------------------------------------------------------------
// A pixel row of such a tensor is [hi(32) | lo(32)] = 64 halves = one 128-byte row, so a 64-channel A unit does not exist.  Compute
// dW^T instead:  D[n][j] = sum_m dz[m][n] * xrow[m][j],  n = output channel (A operand = dz, MN-major), j = 64 entries of the
// interleaved row of one filter tap (B operand, MN-major).  One MMA against dz_hi yields dz_hi*x_hi (columns 0..31) AND dz_hi*x_lo
// (columns 32..63); the one against dz_lo yields dz_lo*x_hi and the 2^-22 term dz_lo*x_lo; dW[tap*32 + c][n] = D[n][c] + D[n][32 + c]
// is a sum inside one thread's registers.  Tile = 128 dz channels (Cout = 64: the 64-row block is read twice, LBO = 0, and the
// duplicate rows are not stored) x 2 filter taps (N = 128); accumulation group g owns tap g of the pair.
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_il32_umma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int n_ablk = p.Cout >= 128 ? 2 : 1;                        // real 64-channel dz blocks per tile
  const int a_bytes = 2 * n_ablk * WG_TILE_BYTES;                  // planes x blocks
  const int stage_bytes = a_bytes + 2 * WG_TILE_BYTES;             // + two taps of interleaved activation rows
  const uint32_t tiles_end = smem_base + p.stages * stage_bytes;
  unsigned char* aux = smem_gen + (size_t)p.stages * stage_bytes;
  const uint32_t bar_full = tiles_end, bar_empty = tiles_end + 8 * WG_MAX_STAGES;
  const uint32_t bar_pfull = tiles_end + 16 * WG_MAX_STAGES, bar_pempty = bar_pfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 16 * WG_MAX_STAGES + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x;
  const int split = unit % p.splits, tile = unit / p.splits;
  const int tp = tile / p.n_tiles_n, mt = tile - tp * p.n_tiles_n;  // tap pair, 128-channel tile of dz
  const int ntaps = p.n_units;                                      // kh*kw
  const int nblk = (p.M + WG_PIX - 1) / WG_PIX;
  const int pb_begin = (int)((long long)nblk * split / p.splits), pb_end = (int)((long long)nblk * (split + 1) / p.splits);
  const int nstage_total = pb_end - pb_begin;
  const int npart = (nstage_total + p.flush - 1) / p.flush;
  constexpr int tmem_cols = 256;                                    // two 128-column partial buffers

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_x);
    prefetch_tmap(&map_dz);
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_pfull + 8 * b, 1); mbar_init(bar_pempty + 8 * b, 8); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int HoWo = p.Ho * p.Wo;
      int tap_r[2], tap_s[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        int tap = 2 * tp + u;
        if (tap >= ntaps) tap = ntaps - 1;                          // odd tap count: the second half of the last tile is not stored
        tap_r[u] = tap / p.kw; tap_s[u] = tap - tap_r[u] * p.kw;
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        const int m0 = pb * WG_PIX;
        const int img = m0 / HoWo, rem = m0 - img * HoWo;
        const int oh = rem / p.Wo, ow = rem - oh * p.Wo;
        const int bw = ow * p.stride - p.pad, bh = oh * p.stride - p.pad;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t full = bar_full + 8 * stage;
        const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + a_bytes;
        mbar_expect_tx(full, (uint32_t)stage_bytes);
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
          for (int j = 0; j < n_ablk; ++j)
            tma_load_2d(sa + (pl * n_ablk + j) * WG_TILE_BYTES, &map_dz, full, mt * 128 + j * 64, (int)(m0 + pl * p.dz_plane_rows));
#pragma unroll
        for (int u = 0; u < 2; ++u)
          tma_load_im2col_4d(sb + u * WG_TILE_BYTES, &map_x, full, 0, bw, bh, img, (uint16_t)tap_s[u], (uint16_t)tap_r[u]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, true, 128, true, true);
      const uint32_t a_lbo = n_ablk == 2 ? (uint32_t)WG_TILE_BYTES : 0u;      // one block: both 64-row halves read the same tile
      int stage = 0;
      uint32_t phase = 0, pcount = 0;
      for (int i = 0; i < nstage_total; ++i) {
        const int pbuf = pcount & 1;
        const uint32_t tmem_main = tmem_base + pbuf * 128;
        if (i % p.flush == 0) {
          mbar_wait(bar_pempty + 8 * pbuf, ((pcount >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + a_bytes;
        uint32_t written = (i % p.flush == 0) ? 0u : 1u;
        const uint64_t bdesc = make_smem_desc_mn(sb, (uint32_t)WG_TILE_BYTES, 1024u);
#pragma unroll
        for (int pi = 0; pi < 2; ++pi) {                                      // dz_lo first (small products), dz_hi closes the stage
          const uint64_t adesc = make_smem_desc_mn(sa + (1 - pi) * n_ablk * WG_TILE_BYTES, a_lbo, 1024u);
#pragma unroll
          for (int k = 0; k < WG_PIX / UMMA_K; ++k) {
            const uint64_t koff = (uint64_t)((k * UMMA_K * 128) >> 4);
            umma_bf16(tmem_main, adesc + koff, bdesc + koff, idesc, written);
            written = 1;
          }
        }
        umma_commit(bar_empty + 8 * stage);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if ((i + 1) % p.flush == 0 || i + 1 == nstage_total) { umma_commit(bar_pfull + 8 * pbuf); ++pcount; }
      }
    }
  } else {
    const int group = (warp - 2) >> 2;                           // = tap of the pair
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;                        // dz channel within the tile
    const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
    float acc[2][32];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[c][i] = 0.f;
    for (int part = 0; part < npart; ++part) {
      const int pbuf = part & 1;
      mbar_wait(bar_pfull + 8 * pbuf, (part >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_addr + pbuf * 128 + group * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c][i] += __uint_as_float(v[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pempty + 8 * pbuf);
    }
    const int tap = 2 * tp + group;
    const int n = mt * 128 + row;
    if (tap < ntaps && row < n_ablk * 64 && n < p.Cout) {
      const float sc = p.acc_scale * (p.acc_scale_dev ? __ldg(p.acc_scale_dev) : 1.f);
      float* op = p.out + (size_t)split * p.out_split_stride + (size_t)(tap * 32) * p.cout_pad + n;
#pragma unroll
      for (int c = 0; c < 32; ++c) op[(size_t)c * p.cout_pad] = (acc[0][c] + acc[1][c]) * sc;      // hi + lo halves of the interleaved row
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// G[i] = sum_s partial[s][i] in split order (deterministic)
__global__ void wgrad_reduce_kernel(const float4* __restrict__ partial, int splits, size_t n4, size_t split_stride4, float4* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = partial[i];
    for (int s = 1; s < splits; ++s) {
      const float4 b = partial[(size_t)s * split_stride4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    out[i] = a;
  }
}

// planes `plane_stride` elements apart; plane_stride < dst_pitch = plane-interleaved rows ([hi(cols) | lo(cols)] per pixel)
__global__ void split_f16x2_kernel(const float* __restrict__ src, long long rows, int cols, int src_pitch, float scale, const float* scale_dev,
                                   __half* __restrict__ dst, int dst_pitch, long long plane_stride, int* sat_flag) {
  const float sc = scale * (scale_dev ? __ldg(scale_dev) : 1.f);
  const int iter_cols = plane_stride < dst_pitch ? cols : dst_pitch;       // planar: zero-fill the padding columns too
  const long long total = rows * iter_cols;
  int sat = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / iter_cols;
    const int c = (int)(i - r * iter_cols);
    const float v = c < cols ? src[r * src_pitch + c] * sc : 0.f;
    if (fabsf(v) > kF16Max) sat = 1;
    const __half h = __float2half_rn(fminf(fmaxf(v, -kF16Max), kF16Max));
    dst[r * dst_pitch + c] = h;
    dst[plane_stride + r * dst_pitch + c] = __float2half_rn(v - __half2float(h));
  }
  if (sat && sat_flag) atomicOr(sat_flag, YOLO_SAT_ACT_BN);
}

int launch_split_f16x2(const float* src, long long rows, int cols, int src_pitch, float scale, const float* scale_dev, void* dst, int dst_pitch,
                       long long plane_stride, int* sat_flag, cudaStream_t st) {
  const long long total = rows * (plane_stride < dst_pitch ? cols : dst_pitch);
  if (total <= 0) return YOLO_OK;
  int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  split_f16x2_kernel<<<blocks, 256, 0, st>>>(src, rows, cols, src_pitch, scale, scale_dev, static_cast<__half*>(dst), dst_pitch, plane_stride, sat_flag);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

bool wgrad_umma_eligible(int cin, int cout, int kh, int kw, int in_dtype, bool in_interleaved) {
  return cin % 64 == 0 && cout % 64 == 0 && kh == kw && in_dtype == DT_F16X2 && !in_interleaved;
}

bool wgrad_umma_il32_eligible(int cin, int cout, int kh, int kw, int in_dtype, bool in_interleaved, int cpitch, int coff) {
  return cin == 32 && in_interleaved && cpitch == 64 && coff == 0 && (cout == 64 || cout % 128 == 0) && kh == kw && in_dtype == DT_F16X2;
}

int wgrad_umma_plan(WgradPlan& w, void* x_base, int x_plane_n, int H, int W, int cin, int cpitch, int coff, int kh, int kw, int stride, int pad,
                    void* dz_base, long long dz_plane_rows, int cout) {
  w.enabled = false;
  void *ft = nullptr, *fi = nullptr;
  int rc = load_tma_entry_points(&ft, &fi);
  if (rc) return rc;
  auto encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ft);
  auto encode_im2col = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fi);
  if (cpitch % 8 || (reinterpret_cast<uintptr_t>(x_base) & 15) || (reinterpret_cast<uintptr_t>(dz_base) & 15)) return YOLO_OK;
  w.cin = cin; w.cout = cout; w.kh = kh; w.kw = kw; w.stride = stride; w.pad = pad; w.H = H; w.W = W;
  w.Ho = (H + 2 * pad - kh) / stride + 1; w.Wo = (W + 2 * pad - kw) / stride + 1;
  w.in_coff = coff; w.x_plane_n = x_plane_n; w.dz_plane_rows = dz_plane_rows;
  w.bn = cout >= 256 && cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64);
  w.il32 = cin == 32 && cpitch == 64;                       // plane-interleaved input: both planes in one 64-element row, N = images only
  {
    cuuint64_t gdim[4] = {(cuuint64_t)cpitch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(w.il32 ? 1 : 2) * x_plane_n};
    cuuint64_t gstr[3] = {(cuuint64_t)cpitch * 2, (cuuint64_t)W * cpitch * 2, (cuuint64_t)H * W * cpitch * 2};
    int lower[2] = {-pad, -pad};
    int upper[2] = {pad - (kw - 1), pad - (kh - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult cr = encode_im2col(reinterpret_cast<CUtensorMap*>(w.map_x), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x_base, gdim, gstr, lower, upper, 64, WG_PIX,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(YOLO_E_CUDA, "wgrad: cuTensorMapEncodeIm2col(C=%d W=%d H=%d N=%d) failed: %d", cpitch, W, H, 2 * x_plane_n, (int)cr);
  }
  {
    cuuint64_t gd[2] = {(cuuint64_t)cout, (cuuint64_t)(2 * dz_plane_rows)};
    cuuint64_t gs[1] = {(cuuint64_t)cout * 2};
    cuuint32_t bx[2] = {64, (cuuint32_t)WG_PIX};
    cuuint32_t es[2] = {1, 1};
    CUresult cr = encode_tiled(reinterpret_cast<CUtensorMap*>(w.map_dz), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dz_base, gd, gs, bx, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(YOLO_E_CUDA, "wgrad: cuTensorMapEncodeTiled(dz %lldx%d) failed: %d", 2 * dz_plane_rows, cout, (int)cr);
  }
  w.enabled = true;
  return YOLO_OK;
}

static void wgrad_shape(const WgradPlan& w, int batch, int num_sms, int& k_tiles, int& n_tiles_n, int& splits) {
  const int n_units = w.il32 ? w.kh * w.kw : w.kh * w.kw * (w.cin / 64);          // il32: units = filter taps, tiles pair them
  k_tiles = (n_units + 1) / 2;
  n_tiles_n = w.il32 ? (w.cout + 127) / 128 : w.cout / w.bn;
  const int tiles = k_tiles * n_tiles_n;
  const int nblk = (batch * w.Ho * w.Wo + WG_PIX - 1) / WG_PIX;
  // split the pixel range until the grid fills the SMs once, keeping at least 8 pixel blocks per split
  splits = tiles >= num_sms ? 1 : num_sms / tiles;
  if (splits > nblk / 8) splits = nblk / 8;
  if (splits < 1) splits = 1;
}

size_t wgrad_umma_scratch_bytes(const WgradPlan& w, int batch, int cout_pad, int num_sms) {
  if (!w.enabled) return 0;
  int kt, ntn, splits;
  wgrad_shape(w, batch, num_sms, kt, ntn, splits);
  return splits > 1 ? (size_t)splits * w.kh * w.kw * w.cin * cout_pad * 4 : 0;
}

int launch_wgrad_umma(const WgradPlan& w, int batch, float* dW, int cout_pad, float acc_scale, const float* acc_scale_dev, float* scratch,
                      size_t scratch_bytes, int variant, cudaStream_t st) {
  if (!w.enabled) return fail(YOLO_E_STATE, "wgrad: plan not built");
  if (batch > w.x_plane_n) return fail(YOLO_E_SHAPE, "wgrad: batch %d exceeds the tensor map's %d", batch, w.x_plane_n);
  if (cout_pad % 4 || (reinterpret_cast<uintptr_t>(dW) & 15)) return fail(YOLO_E_BADARG, "wgrad: dW rows must be 16-byte aligned");
  int num_sms = 0;
  int rc = device_sm_count(&num_sms);
  if (rc) return rc;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int k_tiles;
  wgrad_shape(w, batch, num_sms, k_tiles, p.n_tiles_n, p.splits);
  p.M = batch * w.Ho * w.Wo;
  p.n_units = w.il32 ? w.kh * w.kw : w.kh * w.kw * (w.cin / 64); p.cin_blocks = w.il32 ? 1 : w.cin / 64; p.kw = w.kw;
  p.Ho = w.Ho; p.Wo = w.Wo; p.stride = w.stride; p.pad = w.pad;
  p.in_coff = w.in_coff; p.x_plane_n = w.x_plane_n; p.dz_plane_rows = w.dz_plane_rows;
  p.BN = w.bn;
  p.K = w.kh * w.kw * w.cin; p.Cout = w.cout; p.cout_pad = cout_pad;
  p.acc_scale = acc_scale; p.acc_scale_dev = acc_scale_dev;
  p.variant = variant;
  const size_t kn = (size_t)p.K * cout_pad;
  if (p.splits > 1) {
    if (!scratch || scratch_bytes < (size_t)p.splits * kn * 4) return fail(YOLO_E_BADARG, "wgrad: scratch too small for %d splits", p.splits);
    p.out = scratch; p.out_split_stride = (long long)kn;
  } else {
    p.out = dW; p.out_split_stride = 0;
  }
  const int stage_bytes = w.il32 ? (2 * (w.cout >= 128 ? 2 : 1) + 2) * WG_TILE_BYTES : (4 + 2 * (p.BN / 64)) * WG_TILE_BYTES;
  const int aux_bytes = 16 * WG_MAX_STAGES + 64 + 64;
  int stages = (WG_SMEM_LIMIT - 1024 - aux_bytes) / stage_bytes;
  if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
  if (stages < 2) return fail(YOLO_E_UNSUPPORTED, "wgrad: tile does not fit two pipeline stages");
  p.stages = stages;
  p.flush = 2;                                                    // 8 full-magnitude MMA additions per TMEM partial, like the forward kernel
  const int smem_bytes = 1024 + stages * stage_bytes + aux_bytes;
  const int grid = k_tiles * p.n_tiles_n * p.splits;
  if (w.il32) {
    rc = ensure_dyn_smem(reinterpret_cast<const void*>(&wgrad_il32_umma_kernel), WG_SMEM_LIMIT);
    if (rc) return rc;
    wgrad_il32_umma_kernel<<<grid, WG_THREADS, smem_bytes, st>>>(*reinterpret_cast<const CUtensorMap*>(w.map_x), *reinterpret_cast<const CUtensorMap*>(w.map_dz), p);
  } else {
    rc = ensure_dyn_smem(reinterpret_cast<const void*>(&wgrad_umma_kernel), WG_SMEM_LIMIT);
    if (rc) return rc;
    wgrad_umma_kernel<<<grid, WG_THREADS, smem_bytes, st>>>(*reinterpret_cast<const CUtensorMap*>(w.map_x), *reinterpret_cast<const CUtensorMap*>(w.map_dz), p);
  }
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  if (p.splits > 1) {
    const size_t n4 = kn / 4;
    int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(scratch), p.splits, n4, n4, reinterpret_cast<float4*>(dW));
    ++g_launches;
    YB_CUDA(cudaGetLastError());
  }
  return YOLO_OK;
}

}  // namespace yb

// Unit-test entry (tests/test_gpu_wgrad.py): dW of one convolution from fp32 device tensors.  x: NHWC (N,H,W,cin), dz: (N*Ho*Wo, cout),
// dW: [kh*kw*cin][cout] (k = tap*cin + c).  Operands are split into fp16 planes here; the kernel is the product kernel.
extern "C" int yolo_debug_wgrad(const float* x, const float* dz, int n, int h, int w, int cin, int cout, int k, int stride, int pad, float* dW,
                                int variant, void* stream) {
  using namespace yb;
  if (!x || !dz || !dW) return fail(YOLO_E_BADARG, "debug_wgrad: null argument");
  const bool il = cin == 32;                                    // plane-interleaved activations -> the transposed kernel
  if (!(il ? wgrad_umma_il32_eligible(cin, cout, k, k, DT_F16X2, true, 64, 0) : wgrad_umma_eligible(cin, cout, k, k, DT_F16X2, false)))
    return fail(YOLO_E_UNSUPPORTED, "debug_wgrad: needs cin %% 64 == 0 and cout %% 64 == 0, or cin == 32 with cout == 64 / a multiple of 128");
  cudaStream_t st = (cudaStream_t)stream;
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  const long long M = (long long)n * ho * wo;
  const long long xrows = (long long)n * h * w;
  const long long dz_plane_rows = M + 128;
  __half *xp = nullptr, *dzp = nullptr;
  float* scratch = nullptr;
  YB_CUDA(cudaMalloc(reinterpret_cast<void**>(&xp), (size_t)2 * xrows * cin * 2));
  YB_CUDA(cudaMalloc(reinterpret_cast<void**>(&dzp), (size_t)2 * dz_plane_rows * cout * 2));
  YB_CUDA(cudaMemsetAsync(dzp, 0, (size_t)2 * dz_plane_rows * cout * 2, st));
  int rc = il ? launch_split_f16x2(x, xrows, 32, 32, 1.f, nullptr, xp, 64, 32, nullptr, st)
              : launch_split_f16x2(x, xrows, cin, cin, 1.f, nullptr, xp, cin, xrows * cin, nullptr, st);
  if (!rc) rc = launch_split_f16x2(dz, M, cout, cout, 1.f, nullptr, dzp, cout, dz_plane_rows * cout, nullptr, st);
  WgradPlan plan;
  if (!rc) rc = wgrad_umma_plan(plan, xp, n, h, w, cin, il ? 64 : cin, 0, k, k, stride, pad, dzp, dz_plane_rows, cout);
  if (!rc && !plan.enabled) rc = fail(YOLO_E_UNSUPPORTED, "debug_wgrad: plan not eligible");
  int sms = 0;
  if (!rc) rc = device_sm_count(&sms);
  size_t sb = 0;
  if (!rc) {
    sb = wgrad_umma_scratch_bytes(plan, n, cout, sms);
    if (sb && cudaMalloc(reinterpret_cast<void**>(&scratch), sb) != cudaSuccess) rc = fail(YOLO_E_OOM, "debug_wgrad: scratch");
  }
  if (!rc) rc = launch_wgrad_umma(plan, n, dW, cout, 1.f, nullptr, scratch, sb, variant, st);
  cudaError_t ce = cudaStreamSynchronize(st);
  cudaFree(xp); cudaFree(dzp); if (scratch) cudaFree(scratch);
  if (!rc && ce != cudaSuccess) rc = fail(YOLO_E_CUDA, "debug_wgrad: %s", cudaGetErrorString(ce));
  return rc;
}
