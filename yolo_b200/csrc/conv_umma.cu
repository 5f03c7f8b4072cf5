// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA: activations through an IM2COL tensor map (one bulk copy gathers the 128 output
// pixels x 64 channels of one filter tap, zero-filling the padding halo), weights through a tiled map.
//
// Replaces cuDNN's Convolution+BatchNorm+LeakyReLU(+elemwise_add) operator chain dispatched by MXNet for
// gluoncv `_conv2d` / DarknetBasicBlockV3 / YOLODetectionBlockV3 (reference: yolo_modules/basic_yolo.py:20-26,
// 118-121; car/utils.py:68-95) - BN, activation, residual add, 2x upsample + concat placement and the
// YOLOOutput transpose are all in this kernel's epilogue.  The training step (train_step.cu) runs its forward and
// its data-gradient convolutions through the same kernel: raw conv output + per-channel BatchNorm statistics
// (warp-shuffle reduction in the epilogue), and fp32 accumulate-into-output for the gradient buffers.
//
// GEMM view: D[M x N] = A[M x K] * B[N x K]^T, M = batch*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin,
// k = (r*kw + s)*Cin + c.  CTA tile 128 x BN, K step 64 (one 128-byte swizzle row of bf16).
//
// Precision modes (template MODE):
//   MODE 0  bf16: one operand plane, fp32 accumulate.
//   MODE 2  "fp16x3": every fp32 operand value is carried as two fp16 planes v = v0 + v1 (22-bit split; v1 unscaled - see
//           common.cuh - and the weights pre-scaled by a power of two so that their v1 stays a normal fp16 number).  Three
//           plane pairs are multiplied: (0,0) = the leading product, (0,1) and (1,0) = corrections of 2^-11 relative size.
//           The dropped (1,1) pair is 2^-22 relative - the accuracy class of "3xTF32" at twice its rate.
//   MODE 1  "bf16x6": every fp32 operand value v is carried as three bf16 planes v = v0 + v1 + v2 (exact
//           24-bit split, written by the producing layer's epilogue / packed once for the weights) and the
//           product is formed from the six plane pairs with i + j <= 2.  Dropped pairs are <= 2^-24 relative:
//           fp32-grade operands at 1/6 of the bf16 rate.  The tensor core's fp32 accumulator TRUNCATES (measured:
//           a round-toward-zero bias of ~1e-8 of |acc| per MMA, i.e. -3.4e-5 relative at K = 9216 when all six
//           pairs share one accumulator - profiles/r1_umma_precision.txt): see "two-level accumulation" below.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2..5 and 6..9 = two
// accumulation/epilogue groups (TMEM partial sums -> fp32 registers -> epilogue -> global); tile kinds and the accumulation
// protocol are described at the kernel.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <cmath>

#include <vector>

#include "conv_umma.cuh"
#include "umma_ptx.cuh"

namespace yb {

constexpr int NUM_THREADS = 320;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 227 * 1024;

struct UmmaParams {
  int M, Cout, BN, n_tiles_n, n_tiles;
  int taps, kw, cin_blocks;                 // k-blocks = taps * cin_blocks
  int bk;                                   // K elements per pipeline stage: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows)
  int Ho, Wo, stride, pad;
  int in_coff;
  int a_plane_n;                            // images per plane in the folded N dimension (= max_batch)
  int b_plane_rows;                         // weight rows per plane (= padded Cout)
  int stages;
  int flush;                                // k-blocks per TMEM partial sum (two-level accumulation)
  int hh_last;                              // wide kinds: partial = 2 k-blocks, correction products of both first
  int dbg_nostore;                          // timing experiments: 1 = skip the epilogue's 16-bit stores, 2 = skip the epilogue (wrong results)
  int dbg_pairs;                            // >0: issue only the first n plane pairs (timing experiments; wrong results)
  int a_tiled;                              // 1x1/s1/p0: activations are a plain [pixels][channels] matrix -> tiled TMA, not im2col
  // split-K tail launches: n_tiles counts work UNITS = (tile, K split); unit u works on tile tile_begin + u / ksplit over the k-blocks
  // [split * nkb / ksplit, (split + 1) * nkb / ksplit) and stores its raw fp32 partial tile to out + split * split_stride, rows
  // counted from row_begin (ksplit = 1, tile_begin = row_begin = 0: the ordinary launch)
  int ksplit, tile_begin, row_begin;
  long long split_stride;
  long long a_plane_rows;                   // pixel rows per plane in that matrix (= max_batch*H*W)
  // epilogue
  float acc_scale;                          // power of two undoing the weight pre-scale (exact)
  const float* acc_scale_dev;               // optional second factor read from device memory (dynamic gradient scale of the training step)
  const float* scale;  const float* shift;  int act;
  const void* res;  int res_dtype, res_cpitch, res_coff;  long long res_plane_stride;   // elements
  void* out;  int out_dtype;  int out_cpitch, out_coff;  long long out_plane_stride;  int upsample2;
  int accum;                                // fp32 output only: out += result (gradient buffers with several producers)
  float* stats;                             // optional [ceil(M/32)][2][Cout] per-warp column sums / sums of squares of the result (BatchNorm statistics)
  int* sat_flag;                            // optional: set to 1 when a value exceeds the fp16 range of the high plane
  float bias_comp;                          // truncation-bias compensation: relative amount added back to the accumulated sum (0 = off)
};

// ---------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------
// Two-level accumulation.  The tensor core's fp32 accumulator truncates (round toward zero: a bias of ~1e-8 of
// |acc| per MMA, measured in profiles/r1_umma_precision.txt), which after thousands of MMAs per output is 10-40x the
// rounding noise of an fp32 FMA chain, and systematic.  So a TMEM accumulator only ever holds a short PARTIAL sum; the
// accumulation warps add each partial into fp32 REGISTERS with round-to-nearest while the MMA warp already fills the other
// TMEM partial buffer.
//   128-column tiles: partial = p.flush k-blocks (8 MMAs) of the leading product; the small correction products of the split
//     precisions (2^-11 / 2^-8 of the result) accumulate in a separate TMEM accumulator over the whole K (their truncation
//     is negligible) that is read once per tile.  The two groups alternate tiles: while one converts/stores tile i from its
//     registers, the other one already accumulates the partials of tile i+1 - the epilogue is fully overlapped.
//   128 x 256 tiles (WIDE): the two 256-column partial buffers fill the TMEM, so all plane pairs accumulate in the partial
//     ("merged"), correction pairs first while the partial is still tiny.  Both groups work on every
//     tile (128 columns each); the MMA warp runs at most two partials into the next tile while they store.
//
// Warps: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..5 = accumulation/epilogue group 0, 6..9 = group 1.
constexpr int GROUP_THREADS = 128;
constexpr int STAGE_OUT_BYTES = 32 * 64;   // store-transpose buffer of one epilogue warp

// Column sums over the 32 lanes of a warp of 32 values per lane, transposing butterfly: 31 shuffles instead of 160;
// lane l returns the sum of v[l] over all lanes.  `v` is clobbered.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, off);
    }
  }
  return v[0];
}

// Tile kinds (template KIND):
//   0  one 128 x BN tile per CTA (BN <= 128), separate correction accumulator, the two groups ping-pong over tiles.
//   5  WIDE: one 128 x 256 tile per CTA, merged accumulation; both groups work on every tile, group g owning columns [128 g, 128 g + 128).
//   6  CTA pair (cta_group::2), 256 x 256: each CTA stages its 128 activation rows and HALF of the 256 weight rows (64 KB per stage ->
//      3 stages), accumulators as in kind 5.  Shared-memory traffic per flop is what paces the main loop (profiles/r1_ncu_summary.md).
//   7  C32I: Cin = 32 with PLANE-INTERLEAVED activations (pixel row = [hi(32) | lo(32)] = one 128-byte row instead of two 64-byte rows,
//      which cost twice as much per byte to land).  One k-block = one filter tap, K' = 64: weight plane X = [w_hi | w_hi] gives hi*hi
//      (k-steps 0,1 -> partial) and lo*hi (k-steps 2,3 -> correction), plane Y = [w_lo | 0] gives hi*lo (k-steps 0,1 -> correction).
template <int MODE, bool OUT_F32, int KIND, bool STATS>
// (168 registers: ten warps put three on a scheduler's 16K-register file)
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ UmmaParams p) {
  constexpr bool PAIR = KIND == 6;
  constexpr bool WIDE = KIND == 5 || KIND == 6;
  constexpr bool C32I = KIND == 7;
  static_assert(KIND == 0 || KIND == 5 || KIND == 6 || KIND == 7, "tile kind");
  static_assert(!C32I || MODE == 2, "interleaved planes exist for the fp16 split only");
  constexpr int NP = MODE == 0 ? 1 : (MODE == 1 ? 3 : 2);          // operand planes
  constexpr int NPA = C32I ? 1 : NP;                               // activation tiles per stage
  constexpr int N_PAIRS = MODE == 0 ? 1 : (MODE == 1 ? 6 : 3);     // plane pairs multiplied per k-step
  // MERGE (wide tiles, whose two 256-column partial buffers fill the TMEM): every plane pair accumulates in the partial
  // buffer, correction products first.  The 128-column kinds keep the separate correction accumulator: measured 7 %
  // faster there (two independent accumulation chains) at the same accuracy.
  constexpr bool MERGE = WIDE;
  constexpr bool HAS_CORR = MODE != 0 && !MERGE;
  const int A_TILE_BYTES = TILE_M * p.bk * 2;
  const int row_bytes = p.bk * 2;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [stages][NPA A tiles | NP B tiles] then barriers, tmem pointer, scale/shift staging (one copy per group)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int b_rows = PAIR ? p.BN / 2 : p.BN;                        // weight rows staged by this CTA
  const int b_tile_bytes = b_rows * p.bk * 2;
  const int stage_bytes = NPA * A_TILE_BYTES + NP * b_tile_bytes;
  const uint32_t tiles_end = smem_base + p.stages * stage_bytes;
  unsigned char* aux = smem_gen + (size_t)p.stages * stage_bytes;
  // barrier layout (8 bytes each): full[8] empty[8] pfull[2] pempty[2] cfull[2] cempty[2]
  const uint32_t bar_full = tiles_end, bar_empty = tiles_end + 8 * MAX_STAGES;
  const uint32_t bar_pfull = tiles_end + 16 * MAX_STAGES, bar_pempty = bar_pfull + 16;
  const uint32_t bar_cfull = bar_pfull + 32, bar_cempty = bar_pfull + 48;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 16 * MAX_STAGES + 64);
  float* s_scale_all = reinterpret_cast<float*>(aux + 16 * MAX_STAGES + 128);      // [group][2][BN]: scale, shift
  // per epilogue warp: 32 rows x 64 bytes (one 32-channel chunk of one plane), 64-byte-swizzled - the transpose buffer of the stores
  unsigned char* s_stage_all = aux + ((16 * MAX_STAGES + 128 + 2 * 2 * p.BN * 4 + 127) & ~127);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const int sched_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;    // persistent scheduling unit (CTA or CTA pair)
  const int sched_n = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // TMEM: two partial buffers for the leading product + two correction buffers (tile parity), each `acc_stride` columns
  const int acc_stride = WIDE ? 256 : (p.BN <= 32 ? 32 : (p.BN <= 64 ? 64 : 128));
  const int tmem_cols = WIDE ? 512 : (HAS_CORR ? 4 * acc_stride : (2 * acc_stride < 32 ? 32 : 2 * acc_stride));
  const int nkb = p.taps * p.cin_blocks / p.ksplit;                // k-blocks per work unit (the host keeps ksplit a divisor)
  const int npart = (nkb + p.flush - 1) / p.flush;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_pfull + 8 * b, 1);
      mbar_init(bar_pempty + 8 * b, (PAIR ? 2 : 1) * (WIDE ? 2 : 1) * GROUP_THREADS / 32);      // PAIR: the leader collects both CTAs' groups; WIDE: both groups drain
      mbar_init(bar_cfull + 8 * b, 1);
      mbar_init(bar_cempty + 8 * b, GROUP_THREADS / 32);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // barrier inits visible to the peer before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int HoWo = p.Ho * p.Wo;
      for (int unit = sched_id; unit < p.n_tiles; unit += sched_n) {
        const int tile = p.tile_begin + unit / p.ksplit, kb0 = (unit % p.ksplit) * nkb;
        const int mt = tile / p.n_tiles_n, nt = tile - mt * p.n_tiles_n;
        const int n0 = nt * p.BN + (PAIR ? (int)cta_rank * b_rows : 0);
        const int m0 = (PAIR ? mt * 2 + (int)cta_rank : mt) * TILE_M;
        const int img = m0 / HoWo;
        const int rem = m0 - img * HoWo;
        const int oh = rem / p.Wo, ow = rem - oh * p.Wo;
        const int bw = ow * p.stride - p.pad, bh = oh * p.stride - p.pad;          // receptive-field origin of the first pixel
        for (int kbl = 0; kbl < nkb; ++kbl) {
          const int kb = kb0 + kbl;
          const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
          const int r = tap / p.kw, s = tap - r * p.kw;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + NPA * A_TILE_BYTES;
          if (PAIR) {
            // both CTAs' bytes complete on the LEADER's full barrier, which the leader arms for 2 x stage_bytes
            const uint32_t full0 = mapa_rank0(bar_full + 8 * stage);
            if (cta_rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2u * (uint32_t)stage_bytes);
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {
              if (p.a_tiled)
                tma2_load_2d(sa + pl * A_TILE_BYTES, &map_a, full0, p.in_coff + cb * p.bk, (int)(m0 + pl * p.a_plane_rows));
              else
                tma2_load_im2col_4d(sa + pl * A_TILE_BYTES, &map_a, full0, p.in_coff + cb * p.bk, bw, bh, img + pl * p.a_plane_n,
                                    (uint16_t)s, (uint16_t)r);
              tma2_load_2d(sb + pl * b_tile_bytes, &map_b, full0, kb * p.bk, n0 + pl * p.b_plane_rows);
            }
          } else {
            const uint32_t full = bar_full + 8 * stage;
            mbar_expect_tx(full, (uint32_t)stage_bytes);
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {
              if (pl < NPA) {
                if (p.a_tiled)
                  tma_load_2d(sa + pl * A_TILE_BYTES, &map_a, full, p.in_coff + cb * p.bk, (int)(m0 + pl * p.a_plane_rows));
                else
                  tma_load_im2col_4d(sa + pl * A_TILE_BYTES, &map_a, full, p.in_coff + cb * p.bk, bw, bh, img + pl * p.a_plane_n,
                                     (uint16_t)s, (uint16_t)r);
              }
              tma_load_2d(sb + pl * b_tile_bytes, &map_b, full, kb * p.bk, n0 + pl * p.b_plane_rows);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0 && cta_rank == 0) {
      const uint32_t idesc = make_idesc(p.BN, MODE == 2, PAIR ? 256 : TILE_M);
      const int ksteps = p.bk / UMMA_K;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      uint32_t pcount = 0;                                                     // partial buffers handed out so far
      // sel: 0 = all plane pairs, 1 = correction pairs only, 2 = leading pair only (hh-last ordering over two k-blocks)
      auto issue_pairs = [&](uint32_t sa_tile, uint32_t sb, uint32_t tmem_main, uint32_t tmem_corr, uint32_t& main_written,
                             uint32_t& corr_written, int sel = 0) {
        if constexpr (C32I) {
          const uint64_t adesc = make_smem_desc(sa_tile, row_bytes);
          const uint64_t bx = make_smem_desc(sb, row_bytes), by = make_smem_desc(sb + b_tile_bytes, row_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {                                      // [hi | lo] x [w_hi | w_hi]
            const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
            if (k < 2) { umma_bf16(tmem_main, adesc + koff, bx + koff, idesc, main_written); main_written = 1; }
            else if (p.dbg_pairs == 0 || p.dbg_pairs > 1) { umma_bf16(tmem_corr, adesc + koff, bx + koff, idesc, corr_written); corr_written = 1; }
          }
#pragma unroll
          for (int k = 0; k < 2; ++k) {                                      // hi x w_lo
            const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
            if (p.dbg_pairs == 0 || p.dbg_pairs > 2) { umma_bf16(tmem_corr, adesc + koff, by + koff, idesc, corr_written); corr_written = 1; }
          }
        } else {
#pragma unroll
        for (int pi = 0; pi < N_PAIRS; ++pi) {
          // merged accumulation: the small correction products go FIRST, while the partial sum is still tiny, so that
          // their additions do not add full-magnitude truncation steps; the leading product closes the partial
          const int pair = MERGE ? N_PAIRS - 1 - pi : pi;
          if (p.dbg_pairs != 0 && pair >= p.dbg_pairs) continue;          // timing experiments (-1: no MMA at all, fill + drain only)
          if ((sel == 1 && pair == 0) || (sel == 2 && pair != 0)) continue;
          // pair 0 = leading product (plane 0 x plane 0) -> partial buffer; the rest -> correction accumulator
          constexpr int PA6[6] = {0, 0, 1, 0, 1, 2}, PB6[6] = {0, 1, 0, 2, 1, 0};
          constexpr int PA3[3] = {0, 0, 1}, PB3[3] = {0, 1, 0};
          const int pa = MODE == 0 ? 0 : (MODE == 1 ? PA6[pair] : PA3[pair]);
          const int pb = MODE == 0 ? 0 : (MODE == 1 ? PB6[pair] : PB3[pair]);
          const uint64_t adesc = make_smem_desc(sa_tile + pa * A_TILE_BYTES, row_bytes);
          const uint64_t bdesc = make_smem_desc(sb + pb * b_tile_bytes, row_bytes);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
            if (pair == 0 || MERGE) {
              if (PAIR) umma2_bf16(tmem_main, adesc + koff, bdesc + koff, idesc, main_written);
              else umma_bf16(tmem_main, adesc + koff, bdesc + koff, idesc, main_written);
              main_written = 1;
            } else {
              umma_bf16(tmem_corr, adesc + koff, bdesc + koff, idesc, corr_written);
              corr_written = 1;
            }
          }
        }
        }
      };
      auto commit = [&](uint32_t bar) { if (PAIR) umma2_commit_mc(bar); else umma_commit(bar); };
      for (int tile = sched_id; tile < p.n_tiles; tile += sched_n, ++it) {
        const int cbuf = it & 1;
        if (HAS_CORR) {
          mbar_wait(bar_cempty + 8 * cbuf, (((uint32_t)it >> 1) & 1) ^ 1);   // correction buffer drained (tile it-2)
          tc_fence_after();
        }
        const uint32_t tmem_corr = tmem_base + (2 + cbuf) * acc_stride;
        uint32_t corr_written = 0, tmem_main = 0, main_written = 0;
        int pbuf = 0;
        if (MERGE && p.hh_last) {
          // hh-last partials: one partial = two k-blocks, the correction products of BOTH first, then the leading
          // products - half the TMEM drains of flush 1 with 8 instead of 16 full-magnitude truncation steps
          for (int kb = 0; kb < nkb; kb += 2) {
            pbuf = pcount & 1;
            mbar_wait(bar_pempty + 8 * pbuf, ((pcount >> 1) & 1) ^ 1);
            tc_fence_after();
            tmem_main = tmem_base + pbuf * acc_stride;
            main_written = 0;
            const int n2 = nkb - kb < 2 ? nkb - kb : 2;
            int sj[2];
            uint32_t pj[2];
            for (int j = 0; j < n2; ++j) {
              sj[j] = stage + j; pj[j] = phase;
              if (sj[j] >= p.stages) { sj[j] -= p.stages; pj[j] ^= 1; }
              mbar_wait(bar_full + 8 * sj[j], pj[j]);
              tc_fence_after();
              const uint32_t sa = smem_base + sj[j] * stage_bytes;
              issue_pairs(sa, sa + NPA * A_TILE_BYTES, tmem_main, tmem_corr, main_written, corr_written, 1);
            }
            for (int j = 0; j < n2; ++j) {
              const uint32_t sa = smem_base + sj[j] * stage_bytes;
              issue_pairs(sa, sa + NPA * A_TILE_BYTES, tmem_main, tmem_corr, main_written, corr_written, 2);
              commit(bar_empty + 8 * sj[j]);
            }
            stage += n2;
            if (stage >= p.stages) { stage -= p.stages; phase ^= 1; }
            commit(bar_pfull + 8 * pbuf);
            ++pcount;
          }
        } else {
          for (int kb = 0; kb < nkb; ++kb) {
            if (kb % p.flush == 0) {                                           // start a new partial sum
              pbuf = pcount & 1;
              mbar_wait(bar_pempty + 8 * pbuf, ((pcount >> 1) & 1) ^ 1);
              tc_fence_after();
              tmem_main = tmem_base + pbuf * acc_stride;
              main_written = 0;
            }
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * stage_bytes;
            issue_pairs(sa, sa + NPA * A_TILE_BYTES, tmem_main, tmem_corr, main_written, corr_written);
            commit(bar_empty + 8 * stage);                                    // frees the smem stage (in both CTAs) when the MMAs retire
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            if ((kb + 1) % p.flush == 0 || kb + 1 == nkb) {                    // partial complete -> accumulation warps
              commit(bar_pfull + 8 * pbuf);
              ++pcount;
            }
          }
        }
        if (HAS_CORR) commit(bar_cfull + 8 * cbuf);
      }
    }
  } else {
    // =========================== accumulation + epilogue groups (warps 2..5 and 6..9) ===========================
    const int group = (warp - 2) >> 2;                     // 0 or 1
    const int et = threadIdx.x - 64 - group * GROUP_THREADS;   // 0..127 within the group
    const int lane_grp = warp & 3;                         // TMEM lane quarter this warp may access
    const int row = lane_grp * 32 + lane;                  // accumulator row = pixel within the tile
    const int HoWo = p.Ho * p.Wo;
    const int gcols = WIDE ? 128 : p.BN;                   // accumulator columns this group owns
    float* s_scale = s_scale_all + group * 2 * gcols;
    const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
    const int nchunks = (gcols + 31) >> 5;
    const float acc_scale = p.acc_scale * (p.acc_scale_dev ? __ldg(p.acc_scale_dev) : 1.f);
    int saturated = 0;
    int it = WIDE ? 0 : group;
    int staged_nt = -1;
    // Accumulation turns.  An mbarrier parity wait can only tell "this phase" from "the previous one", so a group must
    // not start waiting for its partials before the other group has consumed all of the preceding tile's partials:
    // named barriers 3/4 pass the turn (FA3-style ping-pong); group 1 donates the first turn to group 0.
    if (!WIDE && group == 1) asm volatile("bar.arrive 3, 256;" ::: "memory");
    for (int unit = sched_id + (WIDE ? 0 : group * sched_n); unit < p.n_tiles; unit += (WIDE ? 1 : 2) * sched_n, it += (WIDE ? 1 : 2)) {
      const int tile = p.tile_begin + unit / p.ksplit, split = unit % p.ksplit;
      const int mt = tile / p.n_tiles_n, nt = tile - mt * p.n_tiles_n;
      const int m0 = (PAIR ? mt * 2 + (int)cta_rank : mt) * TILE_M;
      const int n0 = nt * p.BN + (WIDE ? group * 128 : 0);
      // stage scale/shift of this tile's columns (the group's previous tile is completely finished here); a group that stays on
      // the same column block keeps what it staged
      if (nt != staged_nt) {
        staged_nt = nt;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
        for (int i = et; i < gcols; i += GROUP_THREADS) {
          const int n = n0 + i;
          // the accumulator scale is a power of two (weight pre-scale, dynamic gradient scale): folding it into the BN scale is exact
          s_scale[i] = ((p.scale && n < p.Cout) ? __ldg(p.scale + n) : 1.f) * acc_scale;
          s_scale[gcols + i] = (p.shift && n < p.Cout) ? __ldg(p.shift + n) : 0.f;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
      }

      float acc[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c][i] = 0.f;

      // ---- level 2: add the TMEM partial sums into registers (round-to-nearest) ----
      if (!WIDE) asm volatile("bar.sync %0, 256;" ::"r"(3 + group) : "memory");           // my turn
      uint32_t pc = (uint32_t)it * (uint32_t)npart;
      for (int part = 0; part < npart; ++part, ++pc) {
        const int pbuf = (int)(pc & 1);
        mbar_wait(bar_pfull + 8 * pbuf, (pc >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_addr + pbuf * acc_stride + (WIDE ? group * 128 : 0);
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
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(mapa_rank0(bar_pempty + 8 * pbuf)); else mbar_arrive(bar_pempty + 8 * pbuf); }
      }
      if (HAS_CORR) {
        const int cbuf = it & 1;
        mbar_wait(bar_cfull + 8 * cbuf, ((uint32_t)it >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_addr + (2 + cbuf) * acc_stride;
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
        if (lane == 0) mbar_arrive(bar_cempty + 8 * cbuf);
      }

      if (!WIDE) asm volatile("bar.arrive %0, 256;" ::"r"(4 - group) : "memory");         // the other group's turn

      // ---- epilogue from registers: BN affine, activation, residual, format split, store ----
      // ONE copy of the 32-column code, executed nchunks times on acc[0] with the other chunks rotated down through registers.
      // Unrolled over the four chunks it was 4500 straight-line SASS instructions per tile and spent half of its time waiting
      // for instruction fetch (ncu: stall_no_inst 48 % of the epilogue samples, ~15 us per tile - exposed on the wide kinds,
      // where both groups work on the same tile, and the whole cost of the short-K 1x1 layers; profiles/r2_ncu_conv_1x1.txt).
      const int m = m0 + row;
      const bool valid = m < p.M && p.dbg_nostore != 2;
      // 16-bit outputs leave through a per-warp transpose: a thread owns one pixel row (64 contiguous bytes per plane and chunk), so
      // a direct STG.128 touches 32 different lines per instruction and the LSU pays 32 tag cycles for 512 bytes (measured: the
      // stores were 70 % of the epilogue, 6.7 us per 256 x 256 tile).  Staged through shared memory, lane l stores 16 bytes of row
      // (l >> 2) + 8 i: 8 lines per instruction, whole 32-byte sectors.  Rows are addressed by their pixel index, computed here
      // once per tile for the four rows a lane stores.
      bool rvalid[4];
      int rpix[4];
      {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int mr = m0 + lane_grp * 32 + (lane >> 2) + 8 * i;
          rvalid[i] = mr < p.M && (OUT_F32 ? p.dbg_nostore != 2 : p.dbg_nostore == 0);
          if (p.upsample2) {
            const int img = mr / HoWo, rem = mr - img * HoWo;
            const int oh = rem / p.Wo, ow = rem - oh * p.Wo;
            const int q = p.upsample2 == 1 ? 0 : p.upsample2 - 2;
            rpix[i] = (img * 2 * p.Ho + 2 * oh + (q >> 1)) * (2 * p.Wo) + 2 * ow + (q & 1);
          } else {
            rpix[i] = mr - p.row_begin;
          }
        }
      }
      unsigned char* s_stage = s_stage_all + (warp - 2) * STAGE_OUT_BYTES;
      size_t pix0 = (size_t)(m - p.row_begin);              // fp32 outputs: this thread's own row
      if (OUT_F32 && p.upsample2 > 1) {
        const int img = m / HoWo, rem = m - img * HoWo;
        const int oh = rem / p.Wo, ow = rem - oh * p.Wo;
        const int q = p.upsample2 - 2;
        pix0 = ((size_t)img * 2 * p.Ho + 2 * oh + (q >> 1)) * (2 * p.Wo) + 2 * ow + (q & 1);
      }
      const bool leaky = p.act == ACT_LEAKY, relu = p.act == ACT_RELU;
      float ymax = 0.f;                                     // max |value| this thread stores into a high plane (saturation flag)
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        if (c > 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) { acc[0][i] = acc[1][i]; acc[1][i] = acc[2][i]; acc[2][i] = acc[3][i]; }
        }
        const int c0 = c * 32;
        const int nb = n0 + c0;
        if (nb >= p.Cout) break;
        float* y = acc[0];
        const int nvalid = min(32, p.Cout - nb);
        // fp32 outputs with 16-byte aligned, whole 32-column chunks (every data-gradient buffer) use the same per-warp transpose as the
        // 16-bit formats, 16 columns at a time.  Accumulation (out += result: a gradient range with an earlier producer) is a vector
        // reduction at the L2 (red.global.add.v4.f32): no operand load, no latency, and still deterministic - every element receives
        // exactly one addend per launch, launches are stream-ordered, and a two-operand fp32 add does not depend on order
        const bool f32_tr = OUT_F32 && ((p.out_cpitch | p.out_coff) & 3) == 0 && nvalid == 32;
        {
          const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c0);
          const float4* sh4 = reinterpret_cast<const float4*>(s_scale + gcols + c0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 a = sc4[q], b = sh4[q];
            const float sc[4] = {a.x, a.y, a.z, a.w}, sh[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // The tensor core's accumulator truncates toward zero: a TMEM partial that saw n full-magnitude MMA additions is short by
              // ~n/2 ulp on average, with the SAME sign on every output - a coherent relative bias (measured -1e-7 per layer,
              // profiles/r1_umma_precision.txt) that the next layer's K-sum amplifies.  Every partial carries the same expected
              // relative loss, so it is added back once, here (round-to-nearest of y*(1+c) is unbiased even for c < 1 ulp):
              // Darknet-53 head error vs fp64 2.7e-4 -> 0.9e-4 (profiles/r2_biascomp.txt).
              float t = fmaf(fmaf(y[4 * q + e], p.bias_comp, y[4 * q + e]), sc[e], sh[e]);
              t = leaky ? fmaxf(t, 0.1f * t) : (relu ? fmaxf(t, 0.f) : t);
              y[4 * q + e] = t;
            }
          }
        }
        if constexpr (STATS) {
          // BatchNorm batch statistics of the raw convolution output (training forward): per-warp column sums of the 32 rows
          // by a transposing shuffle butterfly; the finalize kernel adds the per-warp partials in a fixed order (deterministic)
          float t[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) t[i] = valid ? y[i] : 0.f;
          const float s1 = warp_colsum32(t, lane);
#pragma unroll
          for (int i = 0; i < 32; ++i) t[i] = valid ? y[i] * y[i] : 0.f;
          const float s2 = warp_colsum32(t, lane);
          if (lane < nvalid) {
            float* sp = p.stats + ((size_t)(m0 >> 5) + lane_grp) * 2 * p.Cout + nb + lane;
            sp[0] = s1;
            sp[p.Cout] = s2;
          }
        }
        if (OUT_F32) {                                      // head convs: fp32 NHWC == (B, H*W, A, C); gradient buffers (accum)
          float* op = static_cast<float*>(p.out) + (size_t)split * p.split_stride + pix0 * p.out_cpitch + p.out_coff + nb;
          if (f32_tr) {
            const int g = lane & 3, sw = (lane >> 1) & 3;
            float* obase = static_cast<float*>(p.out) + (size_t)split * p.split_stride + p.out_coff + nb + g * 4;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              __syncwarp();
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(s_stage + lane * 64 + ((q ^ sw) << 4)) =
                    make_float4(y[hh * 16 + 4 * q], y[hh * 16 + 4 * q + 1], y[hh * 16 + 4 * q + 2], y[hh * 16 + 4 * q + 3]);
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = (lane >> 2) + 8 * i;
                const float4 v = *reinterpret_cast<const float4*>(s_stage + r * 64 + ((g ^ ((r >> 1) & 3)) << 4));
                float* dst = obase + (size_t)rpix[i] * p.out_cpitch + hh * 16;
                if (!rvalid[i]) continue;
                if (p.accum) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                else *reinterpret_cast<float4*>(dst) = v;
              }
            }
          } else if (!valid) {
            // rows past the last pixel: nothing to store
          } else if (((p.out_cpitch | p.out_coff) & 1) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              if (i + 1 < nvalid) {
                float2 v = make_float2(y[i], y[i + 1]);
                if (p.accum) { const float2 o = *reinterpret_cast<const float2*>(op + i); v.x += o.x; v.y += o.y; }
                *reinterpret_cast<float2*>(op + i) = v;
              } else if (i < nvalid) op[i] = p.accum ? op[i] + y[i] : y[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nvalid) op[i] = p.accum ? op[i] + y[i] : y[i];
          }
        } else {
          constexpr bool F16 = MODE == 2;
          if (p.res && valid) {                             // residual add (DarknetBasicBlockV3), stored in the activation format
            const unsigned short* rp = static_cast<const unsigned short*>(p.res) + (size_t)m * p.res_cpitch + p.res_coff + nb;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {
              const uint4* r4 = reinterpret_cast<const uint4*>(rp + (size_t)pl * p.res_plane_stride);
              uint4 u4[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) u4[q] = __ldg(r4 + q);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t uu[4] = {u4[q].x, u4[q].y, u4[q].z, u4[q].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 f;
                  if (F16) f = __half22float2(*reinterpret_cast<const __half2*>(&uu[e]));
                  else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uu[e]));
                  y[q * 8 + 2 * e] += f.x;
                  y[q * 8 + 2 * e + 1] += f.y;
                }
              }
            }
          }
          // split into the planes of the activation format and store 64 contiguous bytes per plane
#pragma unroll
          for (int pl = 0; pl < NP; ++pl) {
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              if (F16) {
                // the high plane saturates at the fp16 range (cvt.satfinite == clamp to +-65504): flagged below, never silent
                // (yolo_check_saturation).  Columns >= nvalid hold stale TMEM and are never stored - nor tracked.
                if (pl == 0 && i < nvalid && valid) ymax = fmaxf(ymax, fmaxf(fabsf(y[i]), fabsf(y[i + 1])));
                uint32_t h;
                if (pl == 0) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(y[i + 1]), "f"(y[i]));
                else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(y[i + 1]), "f"(y[i]));
                w[i >> 1] = h;
                if (pl == 0) {
                  float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
                  y[i] = y[i] - f.x; y[i + 1] = y[i + 1] - f.y;     // exact: remainder has <= 13 bits
                }
              } else {
                float ha = bf16_round(y[i]), hb = bf16_round(y[i + 1]);
                w[i >> 1] = pack_bf16(ha, hb);
                if (pl + 1 < NP) { y[i] -= ha; y[i + 1] -= hb; }
              }
            }
            // transpose through the warp's staging rows (16-byte chunk g of row r sits at chunk g ^ ((r >> 1) & 3): conflict-free)
            __syncwarp();
            {
              const int sw = (lane >> 1) & 3;
#pragma unroll
              for (int g = 0; g < 4; ++g)
                *reinterpret_cast<uint4*>(s_stage + lane * 64 + ((g ^ sw) << 4)) = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
            }
            __syncwarp();
            {
              const int g = lane & 3;
              unsigned short* obase = static_cast<unsigned short*>(p.out) + (size_t)pl * p.out_plane_stride + p.out_coff + nb + g * 8;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = (lane >> 2) + 8 * i;
                const uint4 v = *reinterpret_cast<const uint4*>(s_stage + r * 64 + ((g ^ ((r >> 1) & 3)) << 4));
                if (rvalid[i] && g * 8 < nvalid) {                              // Cout % 8 == 0
                  *reinterpret_cast<uint4*>(obase + (size_t)rpix[i] * p.out_cpitch) = v;
                  if (p.upsample2 == 1) {                                        // nearest 2x upsampling: three more copies
                    *reinterpret_cast<uint4*>(obase + (size_t)(rpix[i] + 1) * p.out_cpitch) = v;
                    *reinterpret_cast<uint4*>(obase + (size_t)(rpix[i] + 2 * p.Wo) * p.out_cpitch) = v;
                    *reinterpret_cast<uint4*>(obase + (size_t)(rpix[i] + 2 * p.Wo + 1) * p.out_cpitch) = v;
                  }
                }
              }
            }
          }
        }
      }
      if (ymax > kF16Max) saturated = 1;
    }
    if (saturated && p.sat_flag) atomicOr(p.sat_flag, YOLO_SAT_ACT_CONV);
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // the peer may still be arriving on / writing into this CTA
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// Split-K fix-up: sums the fp32 partial tiles of the K splits in a FIXED order (deterministic), applies the BN affine, activation and
// residual, and writes the fp16 planes.  One thread per (row, 8 channels); only tiles of the tail [tile_begin, ...) are touched.
__global__ void __launch_bounds__(256)
splitk_fixup_kernel(const float* __restrict__ ws, int ksplit, long long split_stride, int row_begin, int M, int Cout, int tile_begin,
                    int n_tiles_n, int BN, int rows_per_tile, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                    const unsigned short* __restrict__ res, int res_cpitch, int res_coff, long long res_ps, unsigned short* __restrict__ out,
                    int out_cpitch, int out_coff, long long out_ps, int* sat_flag) {
  const int oct = Cout >> 3;
  const long long total = (long long)(M - row_begin) * oct;
  int sat = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int rl = (int)(i / oct), c = (int)(i - (long long)rl * oct) * 8;
    const int m = row_begin + rl;
    if ((m / rows_per_tile) * n_tiles_n + c / BN < tile_begin) continue;       // this tile belongs to the full-K launch
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = 0.f;
    for (int sp = 0; sp < ksplit; ++sp) {
      const float4* q = reinterpret_cast<const float4*>(ws + (size_t)sp * split_stride + (size_t)rl * Cout + c);
      const float4 a = q[0], b = q[1];
      y[0] += a.x; y[1] += a.y; y[2] += a.z; y[3] += a.w; y[4] += b.x; y[5] += b.y; y[6] += b.z; y[7] += b.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = fmaf(y[j], scale ? __ldg(scale + c + j) : 1.f, shift ? __ldg(shift + c + j) : 0.f);
      t = act == ACT_LEAKY ? fmaxf(t, 0.1f * t) : (act == ACT_RELU ? fmaxf(t, 0.f) : t);
      y[j] = t;
    }
    if (res) {
      const unsigned short* rp = res + (size_t)m * res_cpitch + res_coff + c;
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(rp)), l = __ldg(reinterpret_cast<const uint4*>(rp + res_ps));
      const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hh[e])), b = __half22float2(*reinterpret_cast<const __half2*>(&ll[e]));
        y[2 * e] += a.x + b.x;
        y[2 * e + 1] += a.y + b.y;
      }
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (fmaxf(fabsf(y[2 * e]), fabsf(y[2 * e + 1])) > kF16Max) sat = 1;
      uint32_t h;
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(y[2 * e + 1]), "f"(y[2 * e]));
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
      const __half2 l2 = __floats2half2_rn(y[2 * e] - f.x, y[2 * e + 1] - f.y);
      hi[e] = h;
      lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    unsigned short* op = out + (size_t)m * out_cpitch + out_coff + c;
    *reinterpret_cast<uint4*>(op) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(op + out_ps) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (sat && sat_flag) atomicOr(sat_flag, YOLO_SAT_ACT_CONV);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;
static PFN_cuTensorMapEncodeIm2col_v12000 g_encode_im2col = nullptr;

int load_tma_entry_points(void** tiled, void** im2col) {
  if (!(g_encode_tiled && g_encode_im2col)) {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    YB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(YOLO_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    g_encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    fn = nullptr;
    YB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(YOLO_E_CUDA, "cuTensorMapEncodeIm2col not available from the driver");
    g_encode_im2col = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fn);
  }
  if (tiled) *tiled = reinterpret_cast<void*>(g_encode_tiled);
  if (im2col) *im2col = reinterpret_cast<void*>(g_encode_im2col);
  return YOLO_OK;
}

// Experiment switches are read ONCE per process (they used to be getenv calls on every launch).
struct UmmaEnv {
  bool disable, wide, pairwide, a_tiled, splitk;
  int hhlast;        // -1 default, 0 / 1 forced
  int flush;         // 0 default
  int dbg_pairs, dbg_nostore;
  float bias_comp;   // < 0: per-kind default
};
static const UmmaEnv& umma_env() {
  static const UmmaEnv e = [] {
    auto flag = [](const char* n, bool dflt) { const char* v = getenv(n); return v ? v[0] != '0' : dflt; };
    auto num = [](const char* n, int dflt) { const char* v = getenv(n); return v ? atoi(v) : dflt; };
    UmmaEnv x;
    x.disable = flag("YOLO_B200_DISABLE_UMMA", false);
    x.splitk = flag("YOLO_B200_SPLITK", true);
    x.wide = flag("YOLO_B200_WIDE", true);
    x.pairwide = flag("YOLO_B200_PAIRWIDE", true);
    x.a_tiled = flag("YOLO_B200_A_TILED", true);
    x.hhlast = num("YOLO_B200_HHLAST", -1);
    x.flush = num("YOLO_B200_FLUSH", 0);
    x.dbg_pairs = num("YOLO_B200_DBG_PAIRS", 0);
    x.dbg_nostore = num("YOLO_B200_DBG_NOSTORE", 0);
    { const char* v = getenv("YOLO_B200_BIASCOMP"); x.bias_comp = v ? (float)atof(v) : -1.f; }
    return x;
  }();
  return e;
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
static inline unsigned short f2h(float f) { __half h = __float2half_rn(f); unsigned short u; memcpy(&u, &h, 2); return u; }
static inline float h2f(unsigned short u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

static int pick_bn(int cout) {
  int c16 = (cout + 15) & ~15;
  return c16 < 128 ? c16 : 128;
}
static int mode_of(int precision) { return precision == YOLO_PREC_BF16 ? 0 : (precision == YOLO_PREC_BF16X6 ? 1 : 2); }
static int planes_of(int precision) { return precision == YOLO_PREC_BF16 ? 1 : (precision == YOLO_PREC_BF16X6 ? 3 : 2); }
static int act_dtype_of(int precision) {
  return precision == YOLO_PREC_BF16 ? DT_BF16 : (precision == YOLO_PREC_BF16X6 ? DT_BF16X3 : DT_F16X2);
}

// fp16 planes: scale the weights by a power of two so that max|w| lands in [2^(top-1), 2^top): the low plane (2^-12 of the
// value, stored unscaled) then stays a NORMAL fp16 number for all but negligible weights.  Exact; undone in the epilogue.
// Inference packs with top = 9 ([256, 512)); the training step uses top = 5 so that the weights can grow 2^11-fold before
// the high plane would saturate (the device packer flags that).
static void set_prescale(UmmaConv& u, float wmax, int top) {
  u.prescale = 1.f;
  u.acc_scale = 1.f;
  if (u.precision == YOLO_PREC_FP16X3 && wmax > 0.f && std::isfinite(wmax)) {
    int e;
    frexpf(wmax, &e);                                     // wmax = f * 2^e, f in [0.5, 1)
    int s = top - e;
    s = s < -40 ? -40 : (s > 40 ? 40 : s);
    u.prescale = ldexpf(1.f, s);
    u.acc_scale = ldexpf(1.f, -s);
  }
}

int umma_prepare_weights(UmmaConv& u, int precision, const float* w_oihw, int cout, int cin, int kh, int kw, int stride,
                         int pad, int in_dtype, bool has_prologue, int out_nchw, bool in_interleaved, cudaStream_t st, bool rect_ok) {
  u.eligible = false;
  u.enabled = false;
  u.c32i = false;
  if (precision == YOLO_PREC_FP32) return YOLO_OK;
  if (umma_env().disable) return YOLO_OK;
  // shapes the tensor-core kernel takes; everything else stays on the FFMA kernel
  if (cin % 32 != 0 || has_prologue || out_nchw || (kh != kw && !rect_ok) || in_dtype != act_dtype_of(precision)) return YOLO_OK;
  if (stride < 1 || stride > 8 || pad > 127) return YOLO_OK;
  if (in_interleaved && (precision != YOLO_PREC_FP16X3 || cin != 32)) return YOLO_OK;      // only the C32I kernel reads that layout
  u.c32i = in_interleaved;
  const int np = planes_of(precision);
  u.precision = precision; u.cout = cout; u.cin = cin; u.kh = kh; u.kw = kw; u.stride = stride; u.pad = pad;
  u.bk = (cin % 64 == 0 || u.c32i) ? 64 : 32;
  u.bn_tile = pick_bn(cout);
  const int n_tiles_n = (cout + u.bn_tile - 1) / u.bn_tile;
  const int rows = n_tiles_n * u.bn_tile;                  // zero padded so a weight tile never crosses a plane
  const size_t K = (size_t)kh * kw * (u.c32i ? 64 : cin);     // C32I: K' = 64 per tap ([hi | lo] activation rows)
  u.rows = rows;
  u.kdim = K;
  if (u.w_packed) { cudaFree(u.w_packed); u.w_packed = nullptr; }
  u.w_bytes = (size_t)np * rows * K * 2;
  if (cudaMalloc(&u.w_packed, u.w_bytes) != cudaSuccess) { cudaGetLastError(); return fail(YOLO_E_OOM, "umma: cudaMalloc(%zu) for packed weights failed", u.w_bytes); }
  if (w_oihw) {
    std::vector<unsigned short> host((size_t)np * rows * K, 0);
    float wmax = 0.f;
    for (size_t i = 0; i < (size_t)cout * cin * kh * kw; ++i) wmax = fmaxf(wmax, fabsf(w_oihw[i]));
    set_prescale(u, wmax, 9);
    const float prescale = u.prescale;
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int r = 0; r < kh; ++r)
          for (int s = 0; s < kw; ++s) {
            float v = w_oihw[(((size_t)o * cin + c) * kh + r) * kw + s] * prescale;
            const size_t k = (size_t)(r * kw + s) * cin + c;
            if (u.c32i) {                                     // plane X = [w_hi | w_hi], plane Y = [w_lo | 0]
              const size_t k2 = (size_t)(r * kw + s) * 64 + c;
              unsigned short h0 = f2h(v);
              host[((size_t)0 * rows + o) * K + k2] = h0;
              host[((size_t)0 * rows + o) * K + k2 + 32] = h0;
              host[((size_t)1 * rows + o) * K + k2] = f2h(v - h2f(h0));
            } else if (precision == YOLO_PREC_FP16X3) {
              unsigned short h0 = f2h(v);
              host[((size_t)0 * rows + o) * K + k] = h0;
              host[((size_t)1 * rows + o) * K + k] = f2h(v - h2f(h0));
            } else {
              for (int pl = 0; pl < np; ++pl) {
                unsigned short h = f2bf(v);
                host[((size_t)pl * rows + o) * K + k] = h;
                v -= bf2f(h);
              }
            }
          }
    YB_CUDA(cudaMemcpyAsync(u.w_packed, host.data(), u.w_bytes, cudaMemcpyHostToDevice, st));
    YB_CUDA(cudaStreamSynchronize(st));
  } else {
    YB_CUDA(cudaMemsetAsync(u.w_packed, 0, u.w_bytes, st));       // padding rows / the zero half of the C32I plane Y stay zero
    u.prescale = 1.f; u.acc_scale = 1.f;
  }
  int rc = load_tma_entry_points(nullptr, nullptr);
  if (rc) return rc;
  const CUtensorMapDataType dt = precision == YOLO_PREC_FP16X3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = u.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)np * rows};
  cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)u.bk, (cuuint32_t)u.bn_tile};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = g_encode_tiled(reinterpret_cast<CUtensorMap*>(u.map_b), dt, 2, u.w_packed, gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(YOLO_E_CUDA, "cuTensorMapEncodeTiled(weights %dx%zu) failed: %d", np * rows, K, (int)cr);
  u.has_map_bw = false;
  if (u.bk == 64 && cout % 256 == 0 && precision != YOLO_PREC_BF16X6 && !u.c32i) {      // 256-row box for the 128 x 256 tiles
    cuuint32_t boxw[2] = {(cuuint32_t)u.bk, 256u};
    cr = g_encode_tiled(reinterpret_cast<CUtensorMap*>(u.map_bw), dt, 2, u.w_packed, gdim, gstr, boxw, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    u.has_map_bw = cr == CUDA_SUCCESS;
  }
  u.eligible = true;
  return YOLO_OK;
}

// ---- device-side weight packing (training: the fp32 master weights change every step) -------------------------------
// w_mat is the [kh*kw*Cin][cout_pad] fp32 matrix of the flat parameter buffer (k = tap*Cin + c).  Forward layout: packed[pl][o][k]
// (a 32 x 32 shared-memory transpose keeps both sides coalesced); data-gradient layout: the dgrad convolution has Cin' = Cout,
// Cout' = Cin and a flipped filter, packed[pl][c][tap'*Cout + o] with tap' = (kh-1-r)*kw + (kw-1-s) - `o` stays the fastest index.
__device__ __forceinline__ void split_f16(float v, unsigned short& hi, unsigned short& lo, int& sat) {
  if (fabsf(v) > kF16Max) sat = 1;
  const float c = fminf(fmaxf(v, -kF16Max), kF16Max);
  const __half h = __float2half_rn(c);
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(__float2half_rn(v - __half2float(h)));
}
// One launch packs every direction of a layer: the 32 x 32 tile of w_mat is written straight into the data-gradient layout(s)
// (`o` fastest on both sides) and through a shared-memory transpose into the forward layout.
//
// Parity classes of the data gradient of a 3x3 / stride 2 / pad 1 convolution: the input pixels (2a + py, 2b + px) see only the taps
// whose row has the parity of py + 1 (py = 0: r = 1 at dz row a; py = 1: r = 2 at row a, r = 0 at row a + 1), so each class is a
// dense kh' x kw' convolution (kh' = 1 + py, kw' = 1 + px) over dz itself - no zero-dilated copy, a quarter of the MMAs.  Every tap
// belongs to exactly one class; packed[pl][c][tap'*Cout + o] with tap' = r'*kw' + s'.
struct PackArgs {
  const float* w; int K, Cin, Cout, cout_pad, kh, kw;
  float ps_fwd, ps_dg;
  unsigned short* out_fwd; int rows_fwd, c32i;
  int dg_mode;                       // 0: none, 1: flipped filter, 2: parity classes
  unsigned short* out_dg[4]; int rows_dg;
  int* sat_flag;
};
__device__ __forceinline__ void split_f16_pair(float v0, float v1, uint32_t& hi, uint32_t& lo, int& sat) {
  unsigned short h0, l0, h1, l1;
  split_f16(v0, h0, l0, sat);
  split_f16(v1, h1, l1, sat);
  hi = (uint32_t)h0 | ((uint32_t)h1 << 16);
  lo = (uint32_t)l0 | ((uint32_t)l1 << 16);
}
// 64 (k) x 64 (o) tiles, two elements per thread on both sides (4-byte stores; every tensor-core shape has even K and Cout)
__global__ void __launch_bounds__(256) pack_weights_kernel(const PackArgs a) {
  __shared__ float tile[64][65];
  const int k0 = blockIdx.x * 64, o0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
  int sat = 0;
  for (int j = ty; j < 64; j += 8) {
    const int k = k0 + j, o = o0 + 2 * tx;
    const bool in = k < a.K && o < a.Cout;                            // Cout is even: o + 1 is inside as well
    const float2 v = in ? *reinterpret_cast<const float2*>(a.w + (size_t)k * a.cout_pad + o) : make_float2(0.f, 0.f);
    tile[j][2 * tx] = v.x;
    tile[j][2 * tx + 1] = v.y;
    if (a.dg_mode && in) {
      const int tap = k / a.Cin, c = k - tap * a.Cin, r = tap / a.kw, s2 = tap - r * a.kw;
      unsigned short* out;
      size_t Kp;
      int tap2;
      if (a.dg_mode == 1) {
        out = a.out_dg[0];
        Kp = (size_t)a.kh * a.kw * a.Cout;
        tap2 = (a.kh - 1 - r) * a.kw + (a.kw - 1 - s2);
      } else {
        const int py = r != 1, px = s2 != 1, kwp = 1 + px;
        out = a.out_dg[py * 2 + px];
        Kp = (size_t)(1 + py) * kwp * a.Cout;
        tap2 = (r == 0 ? 1 : 0) * kwp + (s2 == 0 ? 1 : 0);
      }
      uint32_t hi, lo;
      split_f16_pair(v.x * a.ps_dg, v.y * a.ps_dg, hi, lo, sat);
      *reinterpret_cast<uint32_t*>(out + (size_t)c * Kp + (size_t)tap2 * a.Cout + o) = hi;
      *reinterpret_cast<uint32_t*>(out + ((size_t)a.rows_dg + c) * Kp + (size_t)tap2 * a.Cout + o) = lo;
    }
  }
  if (a.out_fwd) {
    __syncthreads();
    const size_t Kp = a.c32i ? (size_t)(a.K / a.Cin) * 64 : (size_t)a.K;
    for (int j = ty; j < 64; j += 8) {
      const int o = o0 + j, k = k0 + 2 * tx;
      if (o >= a.Cout || k >= a.K) continue;
      uint32_t hi, lo;
      split_f16_pair(tile[2 * tx][j] * a.ps_fwd, tile[2 * tx + 1][j] * a.ps_fwd, hi, lo, sat);
      if (a.c32i) {
        const int tap = k / 32, c = k - tap * 32;
        const size_t k2 = (size_t)tap * 64 + c;
        *reinterpret_cast<uint32_t*>(a.out_fwd + (size_t)o * Kp + k2) = hi;
        *reinterpret_cast<uint32_t*>(a.out_fwd + (size_t)o * Kp + k2 + 32) = hi;
        *reinterpret_cast<uint32_t*>(a.out_fwd + ((size_t)a.rows_fwd + o) * Kp + k2) = lo;
      } else {
        *reinterpret_cast<uint32_t*>(a.out_fwd + (size_t)o * Kp + k) = hi;
        *reinterpret_cast<uint32_t*>(a.out_fwd + ((size_t)a.rows_fwd + o) * Kp + k) = lo;
      }
    }
  }
  if (sat && a.sat_flag) atomicOr(a.sat_flag, YOLO_SAT_WEIGHT);
}

void umma_set_prescale(UmmaConv& u, float wmax, int top) { set_prescale(u, wmax, top); }
bool umma_disabled() { return umma_env().disable; }

int umma_pack_device(const UmmaConv* fwd, const UmmaConv* dg, int n_dg, const float* w_mat, int w_cin, int w_cout, int cout_pad, int kh, int kw,
                     int* sat_flag, cudaStream_t st) {
  PackArgs a;
  memset(&a, 0, sizeof(a));
  a.w = w_mat; a.K = kh * kw * w_cin; a.Cin = w_cin; a.Cout = w_cout; a.cout_pad = cout_pad; a.kh = kh; a.kw = kw; a.sat_flag = sat_flag;
  if (fwd && fwd->eligible) {
    if (fwd->precision != YOLO_PREC_FP16X3) return fail(YOLO_E_UNSUPPORTED, "umma: device weight packing exists for fp16x3 only");
    a.out_fwd = static_cast<unsigned short*>(fwd->w_packed); a.rows_fwd = fwd->rows; a.c32i = fwd->c32i ? 1 : 0; a.ps_fwd = fwd->prescale;
  }
  if (dg && n_dg > 0 && dg[0].eligible) {
    if (n_dg != 1 && n_dg != 4) return fail(YOLO_E_BADARG, "umma: one data-gradient convolution or four parity classes");
    if (n_dg == 4 && (kh != 3 || kw != 3)) return fail(YOLO_E_BADARG, "umma: parity classes exist for 3x3 filters");
    for (int q = 0; q < n_dg; ++q) {
      if (!dg[q].eligible || dg[q].precision != YOLO_PREC_FP16X3) return fail(YOLO_E_UNSUPPORTED, "umma: device weight packing exists for fp16x3 only");
      a.out_dg[q] = static_cast<unsigned short*>(dg[q].w_packed);
    }
    a.dg_mode = n_dg == 4 ? 2 : 1; a.rows_dg = dg[0].rows; a.ps_dg = dg[0].prescale;
  }
  if (!a.out_fwd && !a.dg_mode) return YOLO_OK;
  if ((a.K | w_cout | cout_pad) & 1) return fail(YOLO_E_UNSUPPORTED, "umma: device weight packing needs even K and Cout");
  dim3 grid((a.K + 63) / 64, (w_cout + 63) / 64);
  pack_weights_kernel<<<grid, 256, 0, st>>>(a);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

int umma_build_maps(UmmaConv& u, void* in_base, int max_batch, int H, int W, int C, int cpitch, int coff) {
  u.enabled = false;
  if (!u.eligible) return YOLO_OK;
  int rc = load_tma_entry_points(nullptr, nullptr);
  if (rc) return rc;
  const int np = u.c32i ? 1 : planes_of(u.precision);       // C32I: both planes sit in one 64-element pixel row
  if (u.c32i && (cpitch != 64 || coff != 0)) return YOLO_OK;
  if (cpitch % 8 != 0 || (reinterpret_cast<uintptr_t>(in_base) & 15)) return YOLO_OK;       // TMA stride/address alignment
  (void)C;
  const CUtensorMapDataType dt = u.precision == YOLO_PREC_FP16X3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = u.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  u.a_tiled = false;
  if (u.kh == 1 && u.kw == 1 && u.stride == 1 && u.pad == 0 && umma_env().a_tiled) {
    // 1x1 convolution: A is the plain row-major matrix [planes*max_batch*H*W][cpitch] -> tiled map
    cuuint64_t rows = (cuuint64_t)np * max_batch * H * W;
    cuuint64_t gd[2] = {(cuuint64_t)cpitch, rows};
    cuuint64_t gs[1] = {(cuuint64_t)cpitch * 2};
    cuuint32_t bx[2] = {(cuuint32_t)u.bk, (cuuint32_t)TILE_M};
    cuuint32_t es[2] = {1, 1};
    CUresult cr2 = g_encode_tiled(reinterpret_cast<CUtensorMap*>(u.map_a), dt, 2, in_base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr2 != CUDA_SUCCESS) return fail(YOLO_E_CUDA, "cuTensorMapEncodeTiled(activations %llux%d) failed: %d", (unsigned long long)rows, cpitch, (int)cr2);
    u.a_tiled = true;
    u.a_plane_rows = (long long)max_batch * H * W;
    u.max_batch = max_batch;
    u.enabled = true;
    return YOLO_OK;
  }
  cuuint64_t gdim[4] = {(cuuint64_t)cpitch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)np * max_batch};
  cuuint64_t gstr[3] = {(cuuint64_t)cpitch * 2, (cuuint64_t)W * cpitch * 2, (cuuint64_t)H * W * cpitch * 2};
  int lower[2] = {-u.pad, -u.pad};
  int upper[2] = {u.pad - (u.kw - 1), u.pad - (u.kh - 1)};
  if (u.pad_high_full) { upper[0] = 0; upper[1] = 0; }     // every pixel is a base pixel; taps past the far edge read zeros
  cuuint32_t estr[4] = {1, (cuuint32_t)u.stride, (cuuint32_t)u.stride, 1};
  CUresult cr = g_encode_im2col(reinterpret_cast<CUtensorMap*>(u.map_a), dt, 4, in_base, gdim, gstr, lower, upper, (cuuint32_t)u.bk, TILE_M, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS)
    return fail(YOLO_E_CUDA, "cuTensorMapEncodeIm2col(C=%d W=%d H=%d N=%d k=%d s=%d p=%d) failed: %d", cpitch, W, H, np * max_batch, u.kw, u.stride, u.pad, (int)cr);
  u.max_batch = max_batch;
  u.enabled = true;
  return YOLO_OK;
}

void umma_release(UmmaConv& u) {
  if (u.split_ws) cudaFree(u.split_ws);
  u.split_ws = nullptr; u.split_ws_bytes = 0;
  if (u.w_packed) cudaFree(u.w_packed);
  u.w_packed = nullptr;
  u.eligible = u.enabled = false;
}

// Per-device launch state: SM count and the dynamic shared-memory opt-in of every kernel instantiation (the attribute is per
// device and per function; the C ABI allows handles on several devices in one process).
constexpr int kMaxDevices = 64;
int device_sm_count(int* sms) {
  static std::atomic<int> cache[kMaxDevices];
  int dev = 0;
  YB_CUDA(cudaGetDevice(&dev));
  int v = (dev >= 0 && dev < kMaxDevices) ? cache[dev].load(std::memory_order_relaxed) : 0;
  if (v == 0) {
    YB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < kMaxDevices) cache[dev].store(v, std::memory_order_relaxed);
  }
  *sms = v;
  return YOLO_OK;
}

template <int MODE, bool OUT_F32, int KIND, bool STATS>
static int launch_mode4(const UmmaConv& u, const UmmaParams& p, int smem_bytes, int num_sms, cudaStream_t st) {
  int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&conv_umma_kernel<MODE, OUT_F32, KIND, STATS>), SMEM_LIMIT);
  if (rc) return rc;
  if (KIND == 6) {
    int pairs = num_sms / 2;
    if (p.n_tiles < pairs) pairs = p.n_tiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    YB_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<MODE, OUT_F32, KIND, STATS>, *reinterpret_cast<const CUtensorMap*>(u.map_a),
                               *reinterpret_cast<const CUtensorMap*>(u.map_b), p));
    ++g_launches;
    return YOLO_OK;
  }
  const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
  conv_umma_kernel<MODE, OUT_F32, KIND, STATS><<<grid, NUM_THREADS, smem_bytes, st>>>(*reinterpret_cast<const CUtensorMap*>(u.map_a),
                                                                                      *reinterpret_cast<const CUtensorMap*>(KIND == 5 ? u.map_bw : u.map_b), p);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}
// BatchNorm statistics in the epilogue exist for the training forward only: fp16x3, 16-bit output
template <int MODE, bool OUT_F32, int KIND>
static int launch_mode3(const UmmaConv& u, const UmmaParams& p, int smem_bytes, int num_sms, cudaStream_t st) {
  if constexpr (MODE == 2 && !OUT_F32) {
    if (p.stats) return launch_mode4<MODE, OUT_F32, KIND, true>(u, p, smem_bytes, num_sms, st);
  }
  if (p.stats) return fail(YOLO_E_UNSUPPORTED, "umma: epilogue statistics need fp16x3 and a 16-bit output");
  return launch_mode4<MODE, OUT_F32, KIND, false>(u, p, smem_bytes, num_sms, st);
}
template <int MODE>
static int launch_mode(const UmmaConv& u, const UmmaParams& p, int kind, int smem_bytes, int num_sms, cudaStream_t st) {
  const bool f32 = p.out_dtype == DT_F32;
  if constexpr (MODE == 2) {
    if (kind == 5) return f32 ? launch_mode3<MODE, true, 5>(u, p, smem_bytes, num_sms, st) : launch_mode3<MODE, false, 5>(u, p, smem_bytes, num_sms, st);
    if (kind == 6) return f32 ? launch_mode3<MODE, true, 6>(u, p, smem_bytes, num_sms, st) : launch_mode3<MODE, false, 6>(u, p, smem_bytes, num_sms, st);
    if (kind == 7) return f32 ? launch_mode3<MODE, true, 7>(u, p, smem_bytes, num_sms, st) : launch_mode3<MODE, false, 7>(u, p, smem_bytes, num_sms, st);
  }
  if constexpr (MODE == 0) {
    if (kind == 5 && !f32) return launch_mode3<MODE, false, 5>(u, p, smem_bytes, num_sms, st);
    if (kind == 6 && !f32) return launch_mode3<MODE, false, 6>(u, p, smem_bytes, num_sms, st);
  }
  return f32 ? launch_mode3<MODE, true, 0>(u, p, smem_bytes, num_sms, st) : launch_mode3<MODE, false, 0>(u, p, smem_bytes, num_sms, st);
}

int launch_conv_umma(const UmmaConv& u, const ConvDesc& d, cudaStream_t st, UmmaExtra* ex) {
  if (!u.enabled) return fail(YOLO_E_STATE, "umma: tensor maps not built");
  if (d.N > u.max_batch) return fail(YOLO_E_SHAPE, "umma: batch %d exceeds the tensor map's %d", d.N, u.max_batch);
  const UmmaEnv& env = umma_env();
  int num_sms = 0;
  int rc = device_sm_count(&num_sms);
  if (rc) return rc;
  const int np = planes_of(u.precision);
  const int mode = mode_of(u.precision);
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.M = d.N * d.Ho * d.Wo;
  p.Cout = d.Cout;
  p.BN = u.bn_tile;
  p.bk = u.bk;
  p.n_tiles_n = (d.Cout + p.BN - 1) / p.BN;
  const int m_tiles = (p.M + TILE_M - 1) / TILE_M;
  p.n_tiles = m_tiles * p.n_tiles_n;
  p.taps = d.kh * d.kw; p.kw = d.kw; p.cin_blocks = u.c32i ? 1 : d.Cin / p.bk;
  p.Ho = d.Ho; p.Wo = d.Wo; p.stride = d.stride; p.pad = d.pad;
  p.in_coff = d.in_coff;
  p.a_plane_n = u.max_batch;
  p.a_tiled = u.a_tiled ? 1 : 0;
  p.a_plane_rows = u.a_plane_rows;
  p.b_plane_rows = p.n_tiles_n * p.BN;
  // Tile kind.  Wide tiles (128 x 256, merged accumulation) where Cout % 256 == 0: the main loop is bound by shared-memory traffic
  // per flop (MMA operand reads + TMA fill), which N = 256 cuts by a quarter - measured 1.36x on the Darknet-53 step.  CTA pairs
  // (256 x 256, each CTA stages half of the weight rows -> a third less fill per SM, three stages) where at least two M tiles
  // exist; they pay off only with two-k-block partials in hh-last order, because every partial hand-off crosses the cluster
  // (3x3 512->1024 @26^2: kind 6 587 -> 503 us; Darknet-53 step 14.9 -> 13.8 ms, head error 2.4e-4 -> 2.7e-4).
  int kind = 0;
  const bool f32_wide_ok = mode == 2;                     // fp32-output wide tiles are instantiated for the fp16 split only
  if (u.has_map_bw && env.wide && (d.out_dtype != DT_F32 || f32_wide_ok)) {
    kind = 5;
    p.BN = 256;
    p.n_tiles_n = d.Cout / 256;
    p.n_tiles = m_tiles * p.n_tiles_n;
    if (m_tiles >= 2 && u.bn_tile == 128 && env.pairwide) { kind = 6; p.n_tiles = ((m_tiles + 1) / 2) * p.n_tiles_n; }
  }
  if (u.c32i) kind = 7;
  const int stage_bytes = kind == 7 ? TILE_M * p.bk * 2 + np * p.BN * p.bk * 2 : np * (TILE_M * p.bk * 2 + (kind == 6 ? p.BN / 2 : p.BN) * p.bk * 2);
  const int aux_bytes = 16 * MAX_STAGES + 128 + 2 * 2 * p.BN * 4 + 64 + 128 + 8 * STAGE_OUT_BYTES;
  // 8 MMAs of the leading product per TMEM partial.  Longer partials are ~3 % faster but their truncation bias is
  // systematic (same sign on every output) and compounds through the layers: flush 8 -> Darknet-53 head error 6.6e-4
  // instead of 2.9e-4 (profiles/r1_parity_report.txt).  Env YOLO_B200_FLUSH overrides for experiments.
  p.flush = p.bk == 64 ? 2 : 4;
  if (kind == 5 || kind == 6) p.flush = 1;                 // merged accumulation: one k-block (12 MMAs) per partial ...
  if (kind == 7) p.flush = 4;                              // two hi*hi MMAs per tap -> 8 per partial
  if ((kind == 5 || kind == 6) && (env.hhlast >= 0 ? env.hhlast == 1 : kind == 6)) { p.hh_last = 1; p.flush = 2; }   // ... or two in hh-last order
  if (env.flush >= 1 && env.flush <= 64) p.flush = env.flush;
  {
    // truncation-bias compensation: expected relative loss of one partial ~ delta * 2^-23 per 8 full-magnitude MMA additions
    // (delta calibrated on Darknet-53, profiles/r2_biascomp.txt; YOLO_B200_BIASCOMP overrides, 0 = off)
    const float delta = env.bias_comp >= 0.f ? env.bias_comp : 1.4f;
    const int hh_per_partial = (kind == 7 ? 2 : p.bk / UMMA_K) * p.flush;
    p.bias_comp = delta * 1.1920929e-7f * (float)hh_per_partial / 8.f;
  }
  p.dbg_pairs = env.dbg_pairs;
  p.dbg_nostore = env.dbg_nostore;
  int stages = (SMEM_LIMIT - 1024 - aux_bytes) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return fail(YOLO_E_UNSUPPORTED, "umma: tile does not fit two pipeline stages");
  p.stages = stages;
  p.acc_scale = u.acc_scale;
  p.scale = d.scale; p.shift = d.shift; p.act = d.act;
  p.res = d.res; p.res_dtype = d.out_dtype; p.res_cpitch = d.res_cpitch; p.res_coff = d.res_coff;
  p.out = d.out; p.out_dtype = d.out_dtype; p.out_cpitch = d.out_cpitch; p.out_coff = d.out_coff; p.upsample2 = d.upsample2;
  p.res_plane_stride = d.res_plane_stride;
  p.out_plane_stride = d.out_plane_stride;
  p.sat_flag = d.sat_flag;
  if (ex) {
    p.acc_scale_dev = ex->acc_scale_dev; p.accum = ex->accum; p.stats = ex->stats;
    ex->stats_groups_out = (size_t)(kind == 6 ? 8 * ((m_tiles + 1) / 2) : 4 * m_tiles);
  }
  if (p.accum && d.out_dtype != DT_F32) return fail(YOLO_E_UNSUPPORTED, "umma: accumulate-into-output is an fp32 epilogue");
  if (d.out_dtype != DT_F32 && (((d.out_cpitch | d.out_coff) & 7) || d.Cout % 8)) return fail(YOLO_E_UNSUPPORTED, "umma: 16-bit output needs 16-byte aligned channel slices and Cout %% 8 == 0");
  if (d.res && ((d.res_cpitch | d.res_coff) & 7)) return fail(YOLO_E_UNSUPPORTED, "umma: residual needs 16-byte aligned channel slices");
  if (d.res && (d.Cout % 32 || d.out_dtype == DT_F32)) return fail(YOLO_E_UNSUPPORTED, "umma: residual needs Cout %% 32 == 0 and a 16-bit activation format");
  const int smem_bytes = 1024 + stages * stage_bytes + aux_bytes;
  p.ksplit = 1;
  // ---- split-K tail (inference, fp16x3, 16-bit output).  A persistent grid of S scheduling units runs tiles / S waves; the last
  // wave of the 13^2 and 26^2 maps is mostly empty (512->1024 @13^2, batch 32: 88 tiles on 74 CTA pairs = two waves for 1.19 waves
  // of work).  The full waves run as before; the tail tiles are split along K into `ksplit` units each (a divisor of the k-block
  // count, tail * ksplit <= S), every unit writes its raw fp32 partial tile into a per-layer scratch, and a small fix-up kernel adds
  // the splits in a fixed order and does the epilogue.  Batch-1 frames (every layer is "tail") gain the most.
  if (env.splitk && mode == 2 && !ex && d.out_dtype == DT_F16X2 && !d.upsample2 && p.dbg_nostore == 0 && p.dbg_pairs == 0 && d.Cout % 8 == 0) {
    const int S = kind == 6 ? num_sms / 2 : num_sms;
    const int full = (p.n_tiles / S) * S, tail = p.n_tiles - full;
    const int nkb = p.taps * p.cin_blocks;
    int ksplit = 1;
    if (tail > 0 && 2 * tail <= S)
      for (int k = 2; k <= 16 && tail * k <= S; ++k)
        if (nkb % k == 0 && nkb / k >= 2) ksplit = k;
    const int rows_per_tile = kind == 6 ? 2 * TILE_M : TILE_M;
    if (ksplit > 1) {
      // worth it only where the saved part of the last wave outweighs a second conv launch (~20 us of prologue, pipeline fill and
      // epilogue) plus the fix-up (5-12 us): measured per layer in profiles/r2_splitk.txt - the 1x1 and 104^2/208^2 layers lose
      const double wave_us = 2.0 * rows_per_tile * p.BN * (double)nkb * p.bk * 3.0 / (kind == 6 ? 18.4e6 : 9.2e6);   // MMA time of one full-K tile
      const double gain_us = wave_us * (1.0 - 1.0 / ksplit);
      if (gain_us < (full > 0 ? 30.0 : 8.0)) ksplit = 1;
    }
    if (ksplit > 1) {
      const int row_begin = (full / p.n_tiles_n) * rows_per_tile;
      const size_t rows = (size_t)(p.M - row_begin);
      const size_t need = (size_t)ksplit * rows * d.Cout * sizeof(float);
      if (u.split_ws_bytes < need) {                       // per-layer scratch, grown on first use (warm-up), freed by umma_release
        if (u.split_ws) cudaFree(u.split_ws);
        u.split_ws = nullptr; u.split_ws_bytes = 0;
        if (cudaMalloc(&u.split_ws, need) != cudaSuccess) { cudaGetLastError(); ksplit = 1; }
        else u.split_ws_bytes = need;
      }
      if (ksplit > 1) {
        if (full > 0) {
          UmmaParams pa = p;
          pa.n_tiles = full;
          switch (mode) { default: rc = launch_mode<2>(u, pa, kind, smem_bytes, num_sms, st); }
          if (rc) return rc;
        }
        UmmaParams pb = p;
        pb.n_tiles = tail * ksplit; pb.ksplit = ksplit; pb.tile_begin = full; pb.row_begin = row_begin;
        pb.split_stride = (long long)rows * d.Cout;
        pb.scale = nullptr; pb.shift = nullptr; pb.act = ACT_NONE; pb.res = nullptr;
        pb.out = u.split_ws; pb.out_dtype = DT_F32; pb.out_cpitch = d.Cout; pb.out_coff = 0; pb.sat_flag = nullptr;
        rc = launch_mode<2>(u, pb, kind, smem_bytes, num_sms, st);
        if (rc) return rc;
        const long long total = (long long)rows * (d.Cout / 8);
        const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
        splitk_fixup_kernel<<<blocks, 256, 0, st>>>(static_cast<const float*>(u.split_ws), ksplit, pb.split_stride, row_begin, p.M, d.Cout, full,
                                                    p.n_tiles_n, p.BN, rows_per_tile, d.scale, d.shift, d.act,
                                                    static_cast<const unsigned short*>(d.res), d.res_cpitch, d.res_coff, d.res_plane_stride,
                                                    static_cast<unsigned short*>(d.out), d.out_cpitch, d.out_coff, d.out_plane_stride, d.sat_flag);
        ++g_launches;
        YB_CUDA(cudaGetLastError());
        return YOLO_OK;
      }
    }
  }
  switch (mode) {
    case 0: return launch_mode<0>(u, p, kind, smem_bytes, num_sms, st);
    case 1: return launch_mode<1>(u, p, kind, smem_bytes, num_sms, st);
    default: return launch_mode<2>(u, p, kind, smem_bytes, num_sms, st);
  }
}

// rows of per-warp statistics the epilogue may write for M output pixels (whole tiles, CTA pairs round M tiles up to even)
size_t umma_stats_groups(int M) { return (size_t)((M + 255) / 256) * 8; }

}  // namespace yb
