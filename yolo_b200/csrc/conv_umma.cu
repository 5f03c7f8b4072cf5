// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA: activations through an IM2COL tensor map (one bulk copy gathers the 128 output
// pixels x 64 channels of one filter tap, zero-filling the padding halo), weights through a tiled map.
//
// Replaces cuDNN's Convolution+BatchNorm+LeakyReLU(+elemwise_add) operator chain dispatched by MXNet for
// gluoncv `_conv2d` / DarknetBasicBlockV3 / YOLODetectionBlockV3 (reference: yolo_modules/basic_yolo.py:20-26,
// 118-121; car/utils.py:68-95) - BN, activation, residual add, 2x upsample + concat placement and the
// YOLOOutput transpose are all in this kernel's epilogue.
//
// GEMM view: D[M x N] = A[M x K] * B[N x K]^T, M = batch*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin,
// k = (r*kw + s)*Cin + c.  CTA tile 128 x BN, K step 64 (one 128-byte swizzle row of bf16).
//
// Precision modes (template NP = operand planes):
//   NP = 1  bf16 operands, fp32 accumulate.
//   NP = 3  "bf16x6": every fp32 operand value v is carried as three bf16 planes v = v0 + v1 + v2 (exact
//           24-bit split, written by the producing layer's epilogue / packed once for the weights) and the
//           product is formed from the six plane pairs with i + j <= 2.  Dropped pairs are <= 2^-24 relative:
//           fp32-grade operands at 1/6 of the bf16 rate.  The tensor core's fp32 accumulator TRUNCATES (measured:
//           a round-toward-zero bias of ~1e-8 of |acc| per MMA, i.e. -3.4e-5 relative at K = 9216 when all six
//           pairs share one accumulator - profiles/r1_umma_precision.txt), so the leading a0*b0 products rotate
//           over three TMEM accumulators by k-block and the five correction pairs go to a fourth; the epilogue
//           adds the four in fp32 round-to-nearest.  That cuts the truncation steps on the leading term 18x.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation),
// warps 2..5 = epilogue (TMEM -> registers -> global).  With NP = 1 the accumulator is double-buffered in
// TMEM so the epilogue of tile i overlaps the main loop of tile i+1; with NP = 3 the four accumulators fill
// the 512 TMEM columns (the main loop is 6x longer, the exposed epilogue is a few percent).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include <vector>

#include "conv_umma.cuh"

namespace yb {

constexpr int TILE_M = 128;
constexpr int BLOCK_K = 64;                 // bf16 elements = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = TILE_M * BLOCK_K * 2;      // 16 KB per plane
constexpr int NUM_THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 227 * 1024;

struct UmmaParams {
  int M, Cout, BN, n_tiles_n, n_tiles;
  int taps, kw, cin_blocks;                 // k-blocks = taps * cin_blocks
  int Ho, Wo, stride, pad;
  int in_coff;
  int a_plane_n;                            // images per plane in the folded N dimension (= max_batch)
  int b_plane_rows;                         // weight rows per plane (= padded Cout)
  int stages;
  // epilogue
  const float* scale;  const float* shift;  int act;
  const void* res;  int res_cpitch, res_coff;  long long res_plane_stride;   // elements
  void* out;  int out_dtype;  int out_cpitch, out_coff;  long long out_plane_stride;  int upsample2;
};

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const void* map, uint32_t bar, int c, int w, int h, int n,
                                                   uint16_t off_w, uint16_t off_h) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (tile rows are 128 bytes; 8-row groups 1024 bytes apart).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);             // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                               // leading byte offset      bits [16,30) (ignored for swizzled K-major; CuTe writes 1)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset       bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                               // layout: SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128.
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ---------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ UmmaParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [stages][NP A tiles | NP B tiles] then barriers, tmem pointer, scale/shift staging
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int b_tile_bytes = p.BN * BLOCK_K * 2;
  const int stage_bytes = NP * (A_TILE_BYTES + b_tile_bytes);
  const uint32_t tiles_end = smem_base + p.stages * stage_bytes;
  unsigned char* aux = smem_gen + (size_t)p.stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux);                 // full[8] empty[8] tfull[2] tempty[2]
  const uint32_t bar_full = tiles_end, bar_empty = tiles_end + 8 * MAX_STAGES;
  const uint32_t bar_tfull = tiles_end + 16 * MAX_STAGES, bar_tempty = bar_tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 16 * MAX_STAGES + 32);
  float* s_scale = reinterpret_cast<float*>(aux + 16 * MAX_STAGES + 64);      // [2][BN]: scale, shift
  (void)bars;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // TMEM: NP == 1: two buffers of one accumulator; NP == 3: one buffer of four accumulators (3 main + 1 correction)
  constexpr int N_ACC = NP == 1 ? 1 : 4;
  constexpr int N_BUF = NP == 1 ? 2 : 1;
  const int acc_stride = p.BN <= 32 ? 32 : (p.BN <= 64 ? 64 : (p.BN <= 128 ? 128 : 256));
  const int tmem_cols = NP == 1 ? 2 * acc_stride : 512;
  const int nkb = p.taps * p.cin_blocks;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, EPI_THREADS / 32);
    }
    static_assert(N_ACC * N_BUF <= 4, "TMEM budget");
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
      int stage = 0;
      uint32_t phase = 0;
      const int HoWo = p.Ho * p.Wo;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles_n, nt = tile - mt * p.n_tiles_n;
        const int m0 = mt * TILE_M, n0 = nt * p.BN;
        const int img = m0 / HoWo, rem = m0 - img * HoWo;
        const int oh = rem / p.Wo, ow = rem - oh * p.Wo;
        const int bw = ow * p.stride - p.pad, bh = oh * p.stride - p.pad;     // receptive-field origin of the first pixel
        for (int kb = 0; kb < nkb; ++kb) {
          const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
          const int r = tap / p.kw, s = tap - r * p.kw;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, (uint32_t)stage_bytes);
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + NP * A_TILE_BYTES;
#pragma unroll
          for (int pl = 0; pl < NP; ++pl) {
            tma_load_im2col_4d(sa + pl * A_TILE_BYTES, &map_a, full, p.in_coff + cb * BLOCK_K, bw, bh, img + pl * p.a_plane_n,
                               (uint16_t)s, (uint16_t)r);
            tma_load_2d(sb + pl * b_tile_bytes, &map_b, full, kb * BLOCK_K, n0 + pl * p.b_plane_rows);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int buf = N_BUF == 2 ? (it & 1) : 0;
        const uint32_t use = N_BUF == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;
        mbar_wait(bar_tempty + 8 * buf, (use & 1) ^ 1);                       // epilogue has drained this accumulator set
        tc_fence_after();
        const uint32_t tmem_buf = tmem_base + buf * N_ACC * acc_stride;
        uint32_t written = 0;                                                  // accumulators already holding a partial sum
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + NP * A_TILE_BYTES;
          const int main_acc = NP == 1 ? 0 : kb % 3;
#pragma unroll
          for (int pair = 0; pair < (NP == 1 ? 1 : 6); ++pair) {
            int pa, pb, acc;
            if (NP == 1) { pa = 0; pb = 0; acc = 0; }
            else {
              constexpr int PA[6] = {0, 0, 1, 0, 1, 2};
              constexpr int PB[6] = {0, 1, 0, 2, 1, 0};
              pa = PA[pair]; pb = PB[pair];
              acc = pair == 0 ? main_acc : 3;
            }
            const uint32_t tmem_d = tmem_buf + acc * acc_stride;
            const uint64_t adesc = make_smem_desc(sa + pa * A_TILE_BYTES);
            const uint64_t bdesc = make_smem_desc(sb + pb * b_tile_bytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              umma_bf16(tmem_d, adesc + (uint64_t)((k * UMMA_K * 2) >> 4), bdesc + (uint64_t)((k * UMMA_K * 2) >> 4), idesc,
                        (written >> acc) & 1u);
              written |= 1u << acc;
            }
          }
          umma_commit(bar_empty + 8 * stage);                                 // frees the smem stage when the MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * buf);                                     // accumulators complete -> epilogue
      }
    }
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    const int et = threadIdx.x - 64;                       // 0..127
    const int lane_grp = warp & 3;                         // TMEM lane quarter this warp may access
    const int row = lane_grp * 32 + lane;                  // accumulator row = pixel within the tile
    const int HoWo = p.Ho * p.Wo;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / p.n_tiles_n, nt = tile - mt * p.n_tiles_n;
      const int m0 = mt * TILE_M, n0 = nt * p.BN;
      const int buf = N_BUF == 2 ? (it & 1) : 0;
      const uint32_t use = N_BUF == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;
      // stage scale/shift of this tile's columns (previous tile's readers are past their last use: see barrier below)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = et; i < p.BN; i += EPI_THREADS) {
        const int n = n0 + i;
        s_scale[i] = (p.scale && n < p.Cout) ? __ldg(p.scale + n) : 1.f;
        s_scale[p.BN + i] = (p.shift && n < p.Cout) ? __ldg(p.shift + n) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(bar_tfull + 8 * buf, use & 1);
      tc_fence_after();
      const int m = m0 + row;
      const bool m_ok = m < p.M;
      size_t pix[4];
      int npix = 1;
      if (p.upsample2) {
        const int img = m / HoWo, rem = m - img * HoWo;
        const int oh = rem / p.Wo, ow = rem - oh * p.Wo;
        npix = 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) pix[q] = ((size_t)img * 2 * p.Ho + 2 * oh + (q >> 1)) * (2 * p.Wo) + 2 * ow + (q & 1);
      } else {
        pix[0] = (size_t)m;
      }
      const uint32_t taddr_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + buf * N_ACC * acc_stride;
      const int n_main = NP == 1 ? 1 : (nkb < 3 ? nkb : 3);                   // main accumulators that were written
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr_row + c0, v);
        tmem_ld_wait();
        if (NP == 3) {                                                         // ((D0a + D0b) + D0c) + Dcorr, fp32 RN
          uint32_t u[32];
          for (int a = 1; a < n_main; ++a) {
            tmem_ld_32x32b_x32(taddr_row + a * acc_stride + c0, u);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
          }
          tmem_ld_32x32b_x32(taddr_row + 3 * acc_stride + c0, u);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
        }
        if (c0 + 32 >= p.BN) {                              // last chunk read: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
        const int nb = n0 + c0;
        if (!m_ok || nb >= p.Cout) continue;
        float y[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float t = fmaf(__uint_as_float(v[i]), s_scale[c0 + i], s_scale[p.BN + c0 + i]);
          if (p.act == ACT_LEAKY) t = t > 0.f ? t : 0.1f * t;
          else if (p.act == ACT_RELU) t = fmaxf(t, 0.f);
          y[i] = t;
        }
        const int nvalid = min(32, p.Cout - nb);
        if (p.res) {                                        // residual add (DarknetBasicBlockV3), same format as the output
          const __nv_bfloat16* rp = static_cast<const __nv_bfloat16*>(p.res) + (size_t)m * p.res_cpitch + p.res_coff + nb;
#pragma unroll
          for (int pl = 0; pl < NP; ++pl) {
            const uint4* r4 = reinterpret_cast<const uint4*>(rp + (size_t)pl * p.res_plane_stride);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 u = __ldg(r4 + q);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = __bfloat1622float2(h2[e]);
                y[q * 8 + 2 * e] += f.x;
                y[q * 8 + 2 * e + 1] += f.y;
              }
            }
          }
        }
        if (p.out_dtype == DT_F32) {                        // head convs: fp32 NHWC == (B, H*W, A, C)
          float* op = static_cast<float*>(p.out) + pix[0] * p.out_cpitch + p.out_coff + nb;
          if (nvalid == 32 && ((p.out_cpitch | p.out_coff) & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) reinterpret_cast<float4*>(op)[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
          } else if (((p.out_cpitch | p.out_coff) & 1) == 0 && (nvalid & 1) == 0) {
            for (int i = 0; i < nvalid; i += 2) *reinterpret_cast<float2*>(op + i) = make_float2(y[i], y[i + 1]);
          } else {
            for (int i = 0; i < nvalid; ++i) op[i] = y[i];
          }
        } else {
          // bf16 planes: v = v0 + v1 + v2 (exact split); NP == 1 keeps only v0
          uint32_t w[NP][16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a = y[i], b = y[i + 1];
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {
              float ha = bf16_round(a), hb = bf16_round(b);
              w[pl][i >> 1] = pack_bf16(ha, hb);
              a -= ha; b -= hb;
            }
          }
          for (int q = 0; q < npix; ++q) {
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {
              __nv_bfloat16* op = static_cast<__nv_bfloat16*>(p.out) + (size_t)pl * p.out_plane_stride + pix[q] * p.out_cpitch + p.out_coff + nb;
              if (nvalid == 32) {
#pragma unroll
                for (int g = 0; g < 4; ++g) reinterpret_cast<uint4*>(op)[g] = make_uint4(w[pl][4 * g], w[pl][4 * g + 1], w[pl][4 * g + 2], w[pl][4 * g + 3]);
              } else {
                for (int i = 0; i < nvalid; ++i) {
                  uint32_t u = w[pl][i >> 1];
                  reinterpret_cast<unsigned short*>(op)[i] = (i & 1) ? (unsigned short)(u >> 16) : (unsigned short)(u & 0xFFFFu);
                }
              }
            }
          }
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

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;
static PFN_cuTensorMapEncodeIm2col_v12000 g_encode_im2col = nullptr;

static int load_driver_entry_points() {
  if (g_encode_tiled && g_encode_im2col) return YOLO_OK;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  YB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return fail(YOLO_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  g_encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  fn = nullptr;
  YB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return fail(YOLO_E_CUDA, "cuTensorMapEncodeIm2col not available from the driver");
  g_encode_im2col = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fn);
  return YOLO_OK;
}

static inline unsigned short f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (unsigned short)(u >> 16);     // inf / nan
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7FFFu + lsb;                                                          // round to nearest even
  return (unsigned short)(u >> 16);
}
static inline float bf2f(unsigned short h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static int pick_bn(int cout) {
  int c16 = (cout + 15) & ~15;
  return c16 < 128 ? c16 : 128;
}

int umma_prepare_weights(UmmaConv& u, int precision, const float* w_oihw, int cout, int cin, int kh, int kw, int stride,
                         int pad, int in_dtype, bool has_prologue, int out_nchw, cudaStream_t st) {
  u.eligible = false;
  u.enabled = false;
  if (precision == YOLO_PREC_FP32) return YOLO_OK;
  const char* dis = getenv("YOLO_B200_DISABLE_UMMA");
  if (dis && dis[0] == '1') return YOLO_OK;
  // shapes the tensor-core kernel takes; everything else stays on the FFMA kernel
  if (cin % BLOCK_K != 0 || has_prologue || out_nchw || kh != kw || (in_dtype != DT_BF16 && in_dtype != DT_BF16X3)) return YOLO_OK;
  if (stride < 1 || stride > 8 || pad > 127) return YOLO_OK;
  const int np = precision == YOLO_PREC_BF16X6 ? 3 : 1;
  u.precision = precision; u.cout = cout; u.cin = cin; u.kh = kh; u.kw = kw; u.stride = stride; u.pad = pad;
  u.bn_tile = pick_bn(cout);
  const int n_tiles_n = (cout + u.bn_tile - 1) / u.bn_tile;
  const int rows = n_tiles_n * u.bn_tile;                  // zero padded so a weight tile never crosses a plane
  const size_t K = (size_t)kh * kw * cin;
  std::vector<unsigned short> host((size_t)np * rows * K, 0);
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int r = 0; r < kh; ++r)
        for (int s = 0; s < kw; ++s) {
          float v = w_oihw[(((size_t)o * cin + c) * kh + r) * kw + s];
          const size_t k = (size_t)(r * kw + s) * cin + c;
          for (int pl = 0; pl < np; ++pl) {
            unsigned short h = f2bf(v);
            host[((size_t)pl * rows + o) * K + k] = h;
            v -= bf2f(h);
          }
        }
  if (u.w_packed) { cudaFree(u.w_packed); u.w_packed = nullptr; }
  u.w_bytes = host.size() * 2;
  if (cudaMalloc(&u.w_packed, u.w_bytes) != cudaSuccess) { cudaGetLastError(); return fail(YOLO_E_OOM, "umma: cudaMalloc(%zu) for packed weights failed", u.w_bytes); }
  YB_CUDA(cudaMemcpyAsync(u.w_packed, host.data(), u.w_bytes, cudaMemcpyHostToDevice, st));
  YB_CUDA(cudaStreamSynchronize(st));
  int rc = load_driver_entry_points();
  if (rc) return rc;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)np * rows};
  cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)u.bn_tile};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = g_encode_tiled(reinterpret_cast<CUtensorMap*>(u.map_b), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, u.w_packed, gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(YOLO_E_CUDA, "cuTensorMapEncodeTiled(weights %dx%zu) failed: %d", np * rows, K, (int)cr);
  u.eligible = true;
  return YOLO_OK;
}

int umma_build_maps(UmmaConv& u, void* in_base, int max_batch, int H, int W, int C, int cpitch, int coff) {
  u.enabled = false;
  if (!u.eligible) return YOLO_OK;
  int rc = load_driver_entry_points();
  if (rc) return rc;
  const int np = u.precision == YOLO_PREC_BF16X6 ? 3 : 1;
  if (cpitch % 8 != 0 || (reinterpret_cast<uintptr_t>(in_base) & 15)) return YOLO_OK;       // TMA stride/address alignment
  (void)C; (void)coff;
  cuuint64_t gdim[4] = {(cuuint64_t)cpitch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)np * max_batch};
  cuuint64_t gstr[3] = {(cuuint64_t)cpitch * 2, (cuuint64_t)W * cpitch * 2, (cuuint64_t)H * W * cpitch * 2};
  int lower[2] = {-u.pad, -u.pad};
  int upper[2] = {u.pad - (u.kw - 1), u.pad - (u.kh - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)u.stride, (cuuint32_t)u.stride, 1};
  CUresult cr = g_encode_im2col(reinterpret_cast<CUtensorMap*>(u.map_a), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in_base, gdim, gstr, lower, upper,
                                BLOCK_K, TILE_M, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS)
    return fail(YOLO_E_CUDA, "cuTensorMapEncodeIm2col(C=%d W=%d H=%d N=%d k=%d s=%d p=%d) failed: %d", cpitch, W, H, np * max_batch, u.kw, u.stride, u.pad, (int)cr);
  u.max_batch = max_batch;
  u.enabled = true;
  return YOLO_OK;
}

void umma_release(UmmaConv& u) {
  if (u.w_packed) cudaFree(u.w_packed);
  u.w_packed = nullptr;
  u.eligible = u.enabled = false;
}

static int g_num_sms = 0;

template <int NP>
static int launch_np(const UmmaConv& u, const UmmaParams& p, int smem_bytes, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    YB_CUDA(cudaFuncSetAttribute(conv_umma_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_done = true;
  }
  const int grid = p.n_tiles < g_num_sms ? p.n_tiles : g_num_sms;
  conv_umma_kernel<NP><<<grid, NUM_THREADS, smem_bytes, st>>>(*reinterpret_cast<const CUtensorMap*>(u.map_a),
                                                              *reinterpret_cast<const CUtensorMap*>(u.map_b), p);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

int launch_conv_umma(const UmmaConv& u, const ConvDesc& d, cudaStream_t st) {
  if (!u.enabled) return fail(YOLO_E_STATE, "umma: tensor maps not built");
  if (d.N > u.max_batch) return fail(YOLO_E_SHAPE, "umma: batch %d exceeds the tensor map's %d", d.N, u.max_batch);
  if (g_num_sms == 0) {
    int dev = 0;
    YB_CUDA(cudaGetDevice(&dev));
    YB_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int np = u.precision == YOLO_PREC_BF16X6 ? 3 : 1;
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.M = d.N * d.Ho * d.Wo;
  p.Cout = d.Cout;
  p.BN = u.bn_tile;
  p.n_tiles_n = (d.Cout + p.BN - 1) / p.BN;
  p.n_tiles = ((p.M + TILE_M - 1) / TILE_M) * p.n_tiles_n;
  p.taps = d.kh * d.kw; p.kw = d.kw; p.cin_blocks = d.Cin / BLOCK_K;
  p.Ho = d.Ho; p.Wo = d.Wo; p.stride = d.stride; p.pad = d.pad;
  p.in_coff = d.in_coff;
  p.a_plane_n = u.max_batch;
  p.b_plane_rows = p.n_tiles_n * p.BN;
  const int stage_bytes = np * (A_TILE_BYTES + p.BN * BLOCK_K * 2);
  const int aux_bytes = 16 * MAX_STAGES + 64 + 2 * p.BN * 4 + 64;
  int stages = (SMEM_LIMIT - 1024 - aux_bytes) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return fail(YOLO_E_UNSUPPORTED, "umma: tile does not fit two pipeline stages");
  p.stages = stages;
  p.scale = d.scale; p.shift = d.shift; p.act = d.act;
  p.res = d.res; p.res_cpitch = d.res_cpitch; p.res_coff = d.res_coff;
  p.out = d.out; p.out_dtype = d.out_dtype; p.out_cpitch = d.out_cpitch; p.out_coff = d.out_coff; p.upsample2 = d.upsample2;
  p.res_plane_stride = d.res_plane_stride;
  p.out_plane_stride = d.out_plane_stride;
  if (d.out_dtype != DT_F32 && ((d.out_cpitch | d.out_coff) & 7)) return fail(YOLO_E_UNSUPPORTED, "umma: bf16 output needs 16-byte aligned channel slices");
  if (d.res && ((d.res_cpitch | d.res_coff) & 7)) return fail(YOLO_E_UNSUPPORTED, "umma: residual needs 16-byte aligned channel slices");
  if (d.res && d.Cout % 32) return fail(YOLO_E_UNSUPPORTED, "umma: residual needs Cout %% 32 == 0");
  const int smem_bytes = 1024 + stages * stage_bytes + aux_bytes;
  return np == 3 ? launch_np<3>(u, p, smem_bytes, st) : launch_np<1>(u, p, smem_bytes, st);
}

}  // namespace yb
