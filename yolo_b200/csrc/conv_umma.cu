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
#include <cmath>

#include <vector>

#include "conv_umma.cuh"

namespace yb {

constexpr int TILE_M = 128;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 227 * 1024;

struct UmmaParams {
  int M, Cout, BN, n_tiles_n, n_tiles;
  int mt_begin;                             // first M tile of this launch (a layer may be split into a wide and a narrow launch)
  int taps, kw, cin_blocks;                 // k-blocks = taps * cin_blocks
  int bk;                                   // K elements per pipeline stage: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows)
  int Ho, Wo, stride, pad;
  int in_coff;
  int a_plane_n;                            // images per plane in the folded N dimension (= max_batch)
  int b_plane_rows;                         // weight rows per plane (= padded Cout)
  int stages;
  int flush;                                // k-blocks per TMEM partial sum (two-level accumulation)
  int dual;                                 // 1: CTA tile = two M tiles sharing the weight tile; 2: cta_group::2 pair
  int hh_last;                              // merged kinds: partial = 2 k-blocks, correction products of both first (experiment)
  int b_split;                              // wide tiles: fetch the 256 weight rows as two 128-row boxes (experiment)
  int dbg_nostore;                          // timing experiments: 1 = skip the epilogue's 16-bit stores, 2 = skip the epilogue (wrong results)
  int dbg_pairs;                            // >0: issue only the first n plane pairs (timing experiments; wrong results)
  int prefetch;                             // >0: L2-prefetch the operands of k-block kb + prefetch
  int a_tiled;                              // 1x1/s1/p0: activations are a plain [pixels][channels] matrix -> tiled TMA, not im2col
  long long a_plane_rows;                   // pixel rows per plane in that matrix (= max_batch*H*W)
  // epilogue
  float acc_scale;                          // power of two undoing the weight pre-scale (exact)
  const float* scale;  const float* shift;  int act;
  const void* res;  int res_dtype, res_cpitch, res_coff;  long long res_plane_stride;   // elements
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
// ---- 2-CTA (cta_group::2) variants: the pair's leader (cluster rank 0) owns the `full` barriers and issues the MMAs ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t saddr) {          // same smem offset in CTA rank 0 (shared::cluster address)
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(saddr));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const void* map, uint32_t bar_rank0, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar_rank0), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(uint32_t dst, const void* map, uint32_t bar_rank0, int c, int w, int h, int n,
                                                    uint16_t off_w, uint16_t off_h) {
  asm volatile("cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
               ::"r"(dst), "l"(map), "r"(bar_rank0), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {            // arrive on the same barrier offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// multicast variants (KIND 4): one TMA load lands in the shared memory of every CTA of the mask and completes on each one's barrier
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
               ::"r"(dst), "l"(map), "r"(bar), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {          // cta_group::1 MMA, arrive in every CTA of the mask
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

// L2 prefetch of a future k-block's operand tiles (no shared memory needed): turns first-touch DRAM misses into L2 hits
__device__ __forceinline__ void tma_prefetch_2d(const void* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_im2col_4d(const void* map, int c, int w, int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.im2col [%0, {%1, %2, %3, %4}], {%5, %6};"
               ::"l"(map), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major shared-memory matrix descriptor.  Tile rows are `row_bytes` (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B) and
// 8-row groups are 8*row_bytes apart (stride byte offset).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, int row_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);             // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                               // leading byte offset      bits [16,30) (ignored for swizzled K-major; CuTe writes 1)
  d |= (uint64_t)((8 * row_bytes) >> 4) << 32;          // stride byte offset       bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  d |= (uint64_t)(row_bytes == 128 ? 2 : 4) << 61;      // layout: SWIZZLE_128B = 2, SWIZZLE_64B = 4
  return d;
}
// kind::f16 instruction descriptor: (bf16 | fp16) x same -> fp32, both operands K-major, M = 128.
__device__ __forceinline__ uint32_t make_idesc(int n, bool fp16, int m = TILE_M) {
  const uint32_t fmt = fp16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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
//     ("merged"), correction pairs first while the partial is still tiny, one k-block per partial.  Both groups work on every
//     tile (128 columns each); the MMA warp runs at most two partials into the next tile while they store.
//
// Warps: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..5 = accumulation/epilogue group 0, 6..9 = group 1.
constexpr int GROUP_THREADS = 128;

// DUAL: the CTA tile is 256 x BN = two 128-row M tiles that share every weight tile in shared memory (the main loop
// is paced by the TMA ingest rate, ~3 cycles per 128-byte row: sharing B cuts the rows per MMA from 512 to 384).
// Group g then owns M tile g of every pair, with its own partial/correction buffers and barriers.
// PAIR (KIND 2): two CTAs of a cluster form one 256 x BN tile with tcgen05 cta_group::2: each CTA stages its own 128
// rows of A and HALF of the weight tile, the leader issues M=256 MMAs that read both halves, and each CTA keeps the
// accumulators of its own 128 rows in its own TMEM.  Shared-memory traffic per MMA drops from 8 KB to 6 KB and the
// TMA fill from 64 KB to 48 KB per k-block (213 -> 156 B/clk against the 128 B/clk port).
template <int MODE, bool OUT_F32, int KIND>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ UmmaParams p) {
  // KIND 0: one 128 x BN tile per CTA.   KIND 1: two M tiles per CTA sharing the weight tile (halves along M).
  // KIND 2: CTA pair, 256 x BN.          KIND 3: CTA pair, 256 x 2BN: two N tiles sharing the activation tiles (halves along N).
  // "DUAL" = the per-half protocol: epilogue group g owns half g (its own partial/correction buffer and barriers).
  // KIND 4: cluster of two CTAs with ordinary 128-row MMAs whose M tiles share the weight tile: each CTA loads HALF of B and
  //         multicasts it to both (25 % fewer TMA rows per CTA; the TMA row rate is what paces the main loop).
  constexpr bool DUAL = KIND == 1 || KIND == 3, PAIR = KIND == 2 || KIND == 3 || KIND == 6, MCAST = KIND == 4;
  constexpr bool CLUSTER = PAIR || MCAST;
  constexpr bool HALF_M = KIND == 1, HALF_N = KIND == 3;
  // KIND 5 (WIDE): one 128 x 256 tile per CTA (merged accumulation only: the two 256-column partial buffers fill the TMEM);
  //         both accumulation groups work on every tile, group g owning columns [128 g, 128 g + 128).
  // KIND 6: CTA pair, 256 x 256 with cta_group::2 MMAs of N = 256: each CTA stages its 128 activation rows and HALF of the 256
  //         weight rows (64 KB per stage -> 3 stages), accumulators as in KIND 5.
  constexpr bool WIDE = KIND == 5 || KIND == 6;
  // KIND 7 (C32I): Cin = 32 with PLANE-INTERLEAVED activations (pixel row = [hi(32) | lo(32)] = one 128-byte row instead of two
  //         64-byte rows, which cost twice as much per byte to land).  One k-block = one filter tap, K' = 64: weight plane X =
  //         [w_hi | w_hi] gives hi*hi (k-steps 0,1 -> partial) and lo*hi (k-steps 2,3 -> correction), plane Y = [w_lo | 0]
  //         gives hi*lo (k-steps 0,1 -> correction).  Same six MMAs per tap as the planar layout.
  constexpr bool C32I = KIND == 7;
  static_assert(!C32I || MODE == 2, "interleaved planes exist for the fp16 split only");
  constexpr bool SPLIT = DUAL || WIDE;                            // both groups take part in every scheduling unit
  static_assert(!DUAL || MODE != 0, "dual tiles need the correction-buffer TMEM layout");
  constexpr int NMT = HALF_M ? 2 : 1;                              // A (activation) tiles per stage
  constexpr int NBT = HALF_N ? 2 : 1;                              // B (weight) tiles per stage
  constexpr int NP = MODE == 0 ? 1 : (MODE == 1 ? 3 : 2);          // operand planes
  constexpr int NPA = C32I ? 1 : NP;                               // activation tiles per stage and M tile
  constexpr int N_PAIRS = MODE == 0 ? 1 : (MODE == 1 ? 6 : 3);     // plane pairs multiplied per k-step
  // MERGE (wide tiles, whose two 256-column partial buffers fill the TMEM): every plane pair accumulates in the partial
  // buffer, correction products first.  The 128-column kinds keep the separate correction accumulator: measured 7 %
  // faster there (two independent accumulation chains) at the same accuracy.
  constexpr bool MERGE = WIDE || KIND == 8;                       // KIND 8: KIND 0 with merged accumulation (experiment)
  constexpr bool HAS_CORR = MODE != 0 && !MERGE;
  const int A_TILE_BYTES = TILE_M * p.bk * 2;
  const int row_bytes = p.bk * 2;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [stages][NP A tiles | NP B tiles] then barriers, tmem pointer, scale/shift staging (one copy per group)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int b_rows = PAIR ? p.BN / 2 : p.BN;                        // weight rows staged by this CTA
  const int b_tile_bytes = b_rows * p.bk * 2;
  const int stage_bytes = NPA * NMT * A_TILE_BYTES + NP * NBT * b_tile_bytes;
  const uint32_t tiles_end = smem_base + p.stages * stage_bytes;
  unsigned char* aux = smem_gen + (size_t)p.stages * stage_bytes;
  // barrier layout (8 bytes each): full[8] empty[8] pfull[2] pempty[2] cfull[2] cempty[2]
  const uint32_t bar_full = tiles_end, bar_empty = tiles_end + 8 * MAX_STAGES;
  const uint32_t bar_pfull = tiles_end + 16 * MAX_STAGES, bar_pempty = bar_pfull + 16;
  const uint32_t bar_cfull = bar_pfull + 32, bar_cempty = bar_pfull + 48;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 16 * MAX_STAGES + 64);
  float* s_scale_all = reinterpret_cast<float*>(aux + 16 * MAX_STAGES + 128);      // [group][2][BN]: scale, shift

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = CLUSTER ? cluster_ctarank() : 0u;
  const int sched_id = CLUSTER ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;    // persistent scheduling unit (CTA or CTA pair)
  const int sched_n = CLUSTER ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // TMEM: two partial buffers for the leading product + two correction buffers (tile parity), each `acc_stride` columns
  const int acc_stride = WIDE ? 256 : (p.BN <= 32 ? 32 : (p.BN <= 64 ? 64 : 128));
  const int tmem_cols = WIDE ? 512 : (HAS_CORR ? 4 * acc_stride : (2 * acc_stride < 32 ? 32 : 2 * acc_stride));
  const int nkb = p.taps * p.cin_blocks;
  const int npart = (nkb + p.flush - 1) / p.flush;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, MCAST ? 2 : 1);        // MCAST: both CTAs' MMAs must have consumed the stage
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_pfull + 8 * b, 1);
      mbar_init(bar_pempty + 8 * b, (PAIR ? 2 : 1) * (WIDE ? 2 : 1) * GROUP_THREADS / 32);      // PAIR: the leader collects both CTAs' groups; WIDE: both groups drain
      mbar_init(bar_cfull + 8 * b, 1);
      mbar_init(bar_cempty + 8 * b, (PAIR ? 2 : 1) * GROUP_THREADS / 32);
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
  if (CLUSTER) cluster_sync_all(); else __syncthreads();   // barrier inits visible to the peer before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int HoWo = p.Ho * p.Wo;
      for (int tile = sched_id; tile < p.n_tiles; tile += sched_n) {
        const int mt0 = tile / p.n_tiles_n, nt = tile - mt0 * p.n_tiles_n, mt = mt0 + p.mt_begin;
        const int n0 = nt * NBT * p.BN + (PAIR ? (int)cta_rank * b_rows : 0);   // HALF_N: N tile h starts at n0 + h*BN
        const int half_rows = p.BN / 2;                                           // MCAST: weight rows this CTA fetches for both
        int img[NMT], bw[NMT], bh[NMT];
#pragma unroll
        for (int h = 0; h < NMT; ++h) {
          const int m0 = (CLUSTER ? mt * 2 + (int)cta_rank : mt * NMT + h) * TILE_M;
          img[h] = m0 / HoWo;
          const int rem = m0 - img[h] * HoWo;
          const int oh = rem / p.Wo, ow = rem - oh * p.Wo;
          bw[h] = ow * p.stride - p.pad; bh[h] = oh * p.stride - p.pad;          // receptive-field origin of the first pixel
        }
        for (int kb = 0; kb < nkb; ++kb) {
          const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
          const int r = tap / p.kw, s = tap - r * p.kw;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + NMT * NPA * A_TILE_BYTES;
          if (PAIR) {
            // both CTAs' bytes complete on the LEADER's full barrier, which the leader arms for 2 x stage_bytes
            const uint32_t full0 = mapa_rank0(bar_full + 8 * stage);
            if (cta_rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2u * (uint32_t)stage_bytes);
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {
              if (p.a_tiled)
                tma2_load_2d(sa + pl * A_TILE_BYTES, &map_a, full0, p.in_coff + cb * p.bk, (int)((mt * 2 + (int)cta_rank) * TILE_M + pl * p.a_plane_rows));
              else
                tma2_load_im2col_4d(sa + pl * A_TILE_BYTES, &map_a, full0, p.in_coff + cb * p.bk, bw[0], bh[0], img[0] + pl * p.a_plane_n,
                                    (uint16_t)s, (uint16_t)r);
#pragma unroll
              for (int h = 0; h < NBT; ++h)
                tma2_load_2d(sb + (h * NP + pl) * b_tile_bytes, &map_b, full0, kb * p.bk, n0 + h * p.BN + pl * p.b_plane_rows);
            }
          } else {
            if (p.prefetch > 0 && kb + p.prefetch < nkb) {
              const int kb2 = kb + p.prefetch;
              const int tap2 = kb2 / p.cin_blocks, cb2 = kb2 - tap2 * p.cin_blocks, r2 = tap2 / p.kw, s2 = tap2 - r2 * p.kw;
#pragma unroll
              for (int pl = 0; pl < NP; ++pl) {
                if (!p.a_tiled)
                  tma_prefetch_im2col_4d(&map_a, p.in_coff + cb2 * p.bk, bw[0], bh[0], img[0] + pl * p.a_plane_n, (uint16_t)s2, (uint16_t)r2);
                tma_prefetch_2d(&map_b, kb2 * p.bk, n0 + pl * p.b_plane_rows);
              }
            }
            const uint32_t full = bar_full + 8 * stage;
            mbar_expect_tx(full, (uint32_t)stage_bytes);
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {
#pragma unroll
              for (int h = 0; h < NMT; ++h) {
                if (pl >= NPA) break;
                if (p.a_tiled)
                  tma_load_2d(sa + (h * NP + pl) * A_TILE_BYTES, &map_a, full, p.in_coff + cb * p.bk,
                              (int)((CLUSTER ? mt * 2 + (int)cta_rank : mt * NMT + h) * TILE_M + pl * p.a_plane_rows));
                else
                  tma_load_im2col_4d(sa + (h * NP + pl) * A_TILE_BYTES, &map_a, full, p.in_coff + cb * p.bk, bw[h], bh[h],
                                     img[h] + pl * p.a_plane_n, (uint16_t)s, (uint16_t)r);
              }
              if (MCAST)
                tma_load_2d_mc(sb + pl * b_tile_bytes + (int)cta_rank * half_rows * row_bytes, &map_b, full, kb * p.bk,
                               n0 + (int)cta_rank * half_rows + pl * p.b_plane_rows, (uint16_t)3);
              else if (WIDE && p.b_split) {
                tma_load_2d(sb + pl * b_tile_bytes, &map_b, full, kb * p.bk, n0 + pl * p.b_plane_rows);
                tma_load_2d(sb + pl * b_tile_bytes + b_tile_bytes / 2, &map_b, full, kb * p.bk, n0 + 128 + pl * p.b_plane_rows);
              } else
                tma_load_2d(sb + pl * b_tile_bytes, &map_b, full, kb * p.bk, n0 + pl * p.b_plane_rows);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0 && (cta_rank == 0 || MCAST)) {
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
              if (PAIR) umma2_bf16(tmem_corr, adesc + koff, bdesc + koff, idesc, corr_written);
              else umma_bf16(tmem_corr, adesc + koff, bdesc + koff, idesc, corr_written);
              corr_written = 1;
            }
          }
        }
        }
      };
      auto commit = [&](uint32_t bar) { if (PAIR) umma2_commit_mc(bar); else umma_commit(bar); };
      auto commit_stage = [&](uint32_t bar) { if (PAIR) umma2_commit_mc(bar); else if (MCAST) umma_commit_mc(bar, 3); else umma_commit(bar); };
      for (int tile = sched_id; tile < p.n_tiles; tile += sched_n, ++it) {
        if (!DUAL) {
          const int cbuf = it & 1;
          if (HAS_CORR) {
            mbar_wait(bar_cempty + 8 * cbuf, (((uint32_t)it >> 1) & 1) ^ 1);   // correction buffer drained (tile it-2)
            tc_fence_after();
          }
          const uint32_t tmem_corr = tmem_base + (2 + cbuf) * acc_stride;
          uint32_t corr_written = 0, tmem_main = 0, main_written = 0;
          int pbuf = 0;
          if (MERGE && p.hh_last) {
            // EXPERIMENT (YOLO_B200_HHLAST=1): one partial = two k-blocks, the correction products of BOTH first, then the
            // leading products - half the TMEM drains of flush 1 with 8 instead of 16 full-magnitude truncation steps
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
                commit_stage(bar_empty + 8 * sj[j]);
              }
              stage += n2;
              if (stage >= p.stages) { stage -= p.stages; phase ^= 1; }
              commit(bar_pfull + 8 * pbuf);
              ++pcount;
            }
          } else
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
            commit_stage(bar_empty + 8 * stage);                              // frees the smem stage (in both CTAs) when the MMAs retire
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            if ((kb + 1) % p.flush == 0 || kb + 1 == nkb) {                    // partial complete -> accumulation warps
              commit(bar_pfull + 8 * pbuf);
              ++pcount;
            }
          }
          if (HAS_CORR) commit(bar_cfull + 8 * cbuf);
        } else {
          // two M tiles: partial buffer h / correction buffer h belong to M tile h (and to epilogue group h)
          uint32_t corr_written[2] = {0, 0}, main_written[2] = {0, 0};
          for (int kb = 0; kb < nkb; ++kb) {
            const bool new_part = kb % p.flush == 0;
            const bool end_part = (kb + 1) % p.flush == 0 || kb + 1 == nkb;
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * stage_bytes;
            const uint32_t sb = sa + NMT * NP * A_TILE_BYTES;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (kb == 0) {
                mbar_wait(bar_cempty + 8 * h, ((uint32_t)it & 1) ^ 1);        // group h has read the previous pair's correction
                tc_fence_after();
              }
              if (new_part) {
                mbar_wait(bar_pempty + 8 * h, (pcount & 1) ^ 1);              // group h has drained the previous partial
                tc_fence_after();
                main_written[h] = 0;
              }
              issue_pairs(sa + (HALF_M ? h : 0) * NP * A_TILE_BYTES, sb + (HALF_N ? h : 0) * NP * b_tile_bytes, tmem_base + h * acc_stride,
                          tmem_base + (2 + h) * acc_stride, main_written[h], corr_written[h]);
              if (end_part) commit(bar_pfull + 8 * h);
            }
            commit(bar_empty + 8 * stage);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            if (end_part) ++pcount;
          }
          commit(bar_cfull + 8 * 0);
          commit(bar_cfull + 8 * 1);
        }
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
    int it = SPLIT ? 0 : group;
    // Accumulation turns.  An mbarrier parity wait can only tell "this phase" from "the previous one", so a group must
    // not start waiting for its partials before the other group has consumed all of the preceding tile's partials:
    // named barriers 3/4 pass the turn (FA3-style ping-pong); group 1 donates the first turn to group 0.
    if (!SPLIT && group == 1) asm volatile("bar.arrive 3, 256;" ::: "memory");
    for (int tile = sched_id + (SPLIT ? 0 : group * sched_n); tile < p.n_tiles; tile += (SPLIT ? 1 : 2) * sched_n, it += (SPLIT ? 1 : 2)) {
      const int mt0 = tile / p.n_tiles_n, nt = tile - mt0 * p.n_tiles_n, mt = mt0 + p.mt_begin;
      const int m0 = (HALF_M ? mt * 2 + group : (CLUSTER ? mt * 2 + (int)cta_rank : mt)) * TILE_M;
      const int n0 = (HALF_N ? nt * 2 + group : nt) * p.BN + (WIDE ? group * 128 : 0);
      // stage scale/shift of this tile's columns (the group's previous tile is completely finished here)
      asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
      for (int i = et; i < gcols; i += GROUP_THREADS) {
        const int n = n0 + i;
        s_scale[i] = (p.scale && n < p.Cout) ? __ldg(p.scale + n) : 1.f;
        s_scale[gcols + i] = (p.shift && n < p.Cout) ? __ldg(p.shift + n) : 0.f;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");

      float acc[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c][i] = 0.f;

      // ---- level 2: add the TMEM partial sums into registers (round-to-nearest) ----
      if (!SPLIT) asm volatile("bar.sync %0, 256;" ::"r"(3 + group) : "memory");           // my turn
      uint32_t pc = (uint32_t)it * (uint32_t)npart;
      for (int part = 0; part < npart; ++part, ++pc) {
        const int pbuf = DUAL ? group : (int)(pc & 1);
        mbar_wait(bar_pfull + 8 * pbuf, DUAL ? (pc & 1) : ((pc >> 1) & 1));
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
        const int cbuf = DUAL ? group : (it & 1);
        mbar_wait(bar_cfull + 8 * cbuf, DUAL ? ((uint32_t)it & 1) : (((uint32_t)it >> 1) & 1));
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_addr + (2 + cbuf) * acc_stride;
        const float cw = MODE == 2 ? kF16LoScaleInv : 1.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nchunks) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[c][i] = fmaf(__uint_as_float(v[i]), cw, acc[c][i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(mapa_rank0(bar_cempty + 8 * cbuf)); else mbar_arrive(bar_cempty + 8 * cbuf); }
      }

      if (!SPLIT) asm volatile("bar.arrive %0, 256;" ::"r"(4 - group) : "memory");         // the other group's turn

      // ---- epilogue from registers: BN affine, activation, residual, format split, store ----
      // (kept compact on purpose: the first version of this block was 17k SASS instructions and stalled on
      //  instruction fetch - profiles/r1_ncu_summary.md)
      const int m = m0 + row;
      if (m >= p.M || p.dbg_nostore == 2) continue;
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
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int c0 = c * 32;
        const int nb = n0 + c0;
        if (c >= nchunks || nb >= p.Cout) continue;
        float* y = acc[c];
        const int nvalid = min(32, p.Cout - nb);
        {
          const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c0);
          const float4* sh4 = reinterpret_cast<const float4*>(s_scale + gcols + c0);
          const bool leaky = p.act == ACT_LEAKY, relu = p.act == ACT_RELU;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 a = sc4[q], b = sh4[q];
            const float sc[4] = {a.x, a.y, a.z, a.w}, sh[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float t = fmaf(y[4 * q + e] * p.acc_scale, sc[e], sh[e]);
              t = leaky ? fmaxf(t, 0.1f * t) : (relu ? fmaxf(t, 0.f) : t);
              y[4 * q + e] = t;
            }
          }
        }
        if (OUT_F32) {                                      // head convs: fp32 NHWC == (B, H*W, A, C)
          float* op = static_cast<float*>(p.out) + pix[0] * p.out_cpitch + p.out_coff + nb;
          if (((p.out_cpitch | p.out_coff) & 1) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              if (i + 1 < nvalid) *reinterpret_cast<float2*>(op + i) = make_float2(y[i], y[i + 1]);
              else if (i < nvalid) op[i] = y[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nvalid) op[i] = y[i];
          }
        } else {
          constexpr bool F16 = MODE == 2;
          if (p.res) {                                      // residual add (DarknetBasicBlockV3), stored in the activation format
            const unsigned short* rp = static_cast<const unsigned short*>(p.res) + (size_t)m * p.res_cpitch + p.res_coff + nb;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {
              const uint4* r4 = reinterpret_cast<const uint4*>(rp + (size_t)pl * p.res_plane_stride);
              const float pw = (F16 && pl == 1) ? kF16LoScaleInv : 1.f;
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
                  y[q * 8 + 2 * e] = fmaf(f.x, pw, y[q * 8 + 2 * e]);
                  y[q * 8 + 2 * e + 1] = fmaf(f.y, pw, y[q * 8 + 2 * e + 1]);
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
                float a = y[i], b = y[i + 1];
                if (pl == 0) { a = fminf(fmaxf(a, -kF16Max), kF16Max); b = fminf(fmaxf(b, -kF16Max), kF16Max); }
                __half2 h = __floats2half2_rn(a, b);
                w[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
                if (pl == 0) {
                  float2 f = __half22float2(h);
                  y[i] = (y[i] - f.x) * kF16LoScale; y[i + 1] = (y[i + 1] - f.y) * kF16LoScale;     // exact: remainder has <= 13 bits
                }
              } else {
                float ha = bf16_round(y[i]), hb = bf16_round(y[i + 1]);
                w[i >> 1] = pack_bf16(ha, hb);
                if (pl + 1 < NP) { y[i] -= ha; y[i + 1] -= hb; }
              }
            }
            for (int q = 0; q < npix && !p.dbg_nostore; ++q) {
              uint4* op = reinterpret_cast<uint4*>(static_cast<unsigned short*>(p.out) + (size_t)pl * p.out_plane_stride +
                                                   pix[q] * p.out_cpitch + p.out_coff + nb);
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (g * 8 < nvalid) op[g] = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);   // Cout % 8 == 0
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  if (CLUSTER) cluster_sync_all(); else __syncthreads();   // the peer may still be arriving on / writing into this CTA
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
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

int umma_prepare_weights(UmmaConv& u, int precision, const float* w_oihw, int cout, int cin, int kh, int kw, int stride,
                         int pad, int in_dtype, bool has_prologue, int out_nchw, bool in_interleaved, cudaStream_t st) {
  u.eligible = false;
  u.enabled = false;
  u.c32i = false;
  if (precision == YOLO_PREC_FP32) return YOLO_OK;
  const char* dis = getenv("YOLO_B200_DISABLE_UMMA");
  if (dis && dis[0] == '1') return YOLO_OK;
  // shapes the tensor-core kernel takes; everything else stays on the FFMA kernel
  if (cin % 32 != 0 || has_prologue || out_nchw || kh != kw || in_dtype != act_dtype_of(precision)) return YOLO_OK;
  if (stride < 1 || stride > 8 || pad > 127) return YOLO_OK;
  if (in_interleaved && (precision != YOLO_PREC_FP16X3 || cin != 32)) return YOLO_OK;      // only the C32I kernel reads that layout
  u.c32i = in_interleaved;
  const int np = planes_of(precision);
  u.precision = precision; u.cout = cout; u.cin = cin; u.kh = kh; u.kw = kw; u.stride = stride; u.pad = pad;
  u.bk = (cin % 64 == 0 || u.c32i) ? 64 : 32;
  if (const char* be = getenv("YOLO_B200_BK")) { if (atoi(be) == 32 && !u.c32i) u.bk = 32; }      // experiment: more, smaller pipeline stages
  u.bn_tile = pick_bn(cout);
  const int n_tiles_n = (cout + u.bn_tile - 1) / u.bn_tile;
  const int rows = n_tiles_n * u.bn_tile;                  // zero padded so a weight tile never crosses a plane
  const size_t K = (size_t)kh * kw * (u.c32i ? 64 : cin);     // C32I: K' = 64 per tap ([hi | lo] activation rows)
  std::vector<unsigned short> host((size_t)np * rows * K, 0);
  // fp16 planes: scale the weights by a power of two so that max|w| lands in [256, 512): the low plane (2^-12 of the
  // value, stored unscaled) then stays a NORMAL fp16 number for all but negligible weights.  Exact; undone in the epilogue.
  float prescale = 1.f;
  u.acc_scale = 1.f;
  if (precision == YOLO_PREC_FP16X3) {
    float wmax = 0.f;
    for (size_t i = 0; i < (size_t)cout * cin * kh * kw; ++i) wmax = fmaxf(wmax, fabsf(w_oihw[i]));
    if (wmax > 0.f && std::isfinite(wmax)) {
      int e;
      frexpf(wmax, &e);                                     // wmax = f * 2^e, f in [0.5, 1)
      int s = 9 - e;
      s = s < -40 ? -40 : (s > 40 ? 40 : s);
      prescale = ldexpf(1.f, s);
      u.acc_scale = ldexpf(1.f, -s);
    }
  }
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
            host[((size_t)1 * rows + o) * K + k] = f2h((v - h2f(h0)) * kF16LoScale);
          } else {
            for (int pl = 0; pl < np; ++pl) {
              unsigned short h = f2bf(v);
              host[((size_t)pl * rows + o) * K + k] = h;
              v -= bf2f(h);
            }
          }
        }
  if (u.w_packed) { cudaFree(u.w_packed); u.w_packed = nullptr; }
  u.w_bytes = host.size() * 2;
  if (cudaMalloc(&u.w_packed, u.w_bytes) != cudaSuccess) { cudaGetLastError(); return fail(YOLO_E_OOM, "umma: cudaMalloc(%zu) for packed weights failed", u.w_bytes); }
  YB_CUDA(cudaMemcpyAsync(u.w_packed, host.data(), u.w_bytes, cudaMemcpyHostToDevice, st));
  YB_CUDA(cudaStreamSynchronize(st));
  int rc = load_driver_entry_points();
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
  u.has_map_b2 = false;
  if (u.bn_tile % 32 == 0 && !u.c32i) {                    // half-height box for the 2-CTA path (each CTA stages BN/2 weight rows)
    cuuint32_t box2[2] = {(cuuint32_t)u.bk, (cuuint32_t)(u.bn_tile / 2)};
    cr = g_encode_tiled(reinterpret_cast<CUtensorMap*>(u.map_b2), dt, 2, u.w_packed, gdim, gstr, box2, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    u.has_map_b2 = cr == CUDA_SUCCESS;
  }
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

int umma_build_maps(UmmaConv& u, void* in_base, int max_batch, int H, int W, int C, int cpitch, int coff) {
  u.enabled = false;
  if (!u.eligible) return YOLO_OK;
  int rc = load_driver_entry_points();
  if (rc) return rc;
  const int np = u.c32i ? 1 : planes_of(u.precision);       // C32I: both planes sit in one 64-element pixel row
  if (u.c32i && (cpitch != 64 || coff != 0)) return YOLO_OK;
  if (cpitch % 8 != 0 || (reinterpret_cast<uintptr_t>(in_base) & 15)) return YOLO_OK;       // TMA stride/address alignment
  (void)C;
  const CUtensorMapDataType dt = u.precision == YOLO_PREC_FP16X3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = u.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  u.a_tiled = false;
  const char* te = getenv("YOLO_B200_A_TILED");
  if (u.kh == 1 && u.stride == 1 && u.pad == 0 && !(te && te[0] == '0')) {
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
  if (u.w_packed) cudaFree(u.w_packed);
  u.w_packed = nullptr;
  u.eligible = u.enabled = false;
}

static int g_num_sms = 0;

template <int MODE, bool OUT_F32, int KIND>
static int launch_mode3(const UmmaConv& u, const UmmaParams& p, int smem_bytes, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    YB_CUDA(cudaFuncSetAttribute(conv_umma_kernel<MODE, OUT_F32, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_done = true;
  }
  if ((KIND >= 2 && KIND <= 4) || KIND == 6) {
    int pairs = g_num_sms / 2;
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
    YB_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<MODE, OUT_F32, KIND>, *reinterpret_cast<const CUtensorMap*>(u.map_a),
                               *reinterpret_cast<const CUtensorMap*>(KIND == 6 ? u.map_b : u.map_b2), p));
    ++g_launches;
    return YOLO_OK;
  }
  const int grid = p.n_tiles < g_num_sms ? p.n_tiles : g_num_sms;
  conv_umma_kernel<MODE, OUT_F32, KIND><<<grid, NUM_THREADS, smem_bytes, st>>>(*reinterpret_cast<const CUtensorMap*>(u.map_a),
                                                                               *reinterpret_cast<const CUtensorMap*>((KIND == 5 && !p.b_split) ? u.map_bw : u.map_b), p);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}
template <int MODE>
static int launch_mode(const UmmaConv& u, const UmmaParams& p, int smem_bytes, cudaStream_t st) {
  if constexpr (MODE != 1) {
    if (p.dual == 5 && p.out_dtype != DT_F32) return launch_mode3<MODE, false, 5>(u, p, smem_bytes, st);
    if (p.dual == 6 && p.out_dtype != DT_F32) return launch_mode3<MODE, false, 6>(u, p, smem_bytes, st);
  }
  if constexpr (MODE == 2) {
    if (p.dual == 7) return p.out_dtype == DT_F32 ? launch_mode3<MODE, true, 7>(u, p, smem_bytes, st) : launch_mode3<MODE, false, 7>(u, p, smem_bytes, st);
    if (p.dual == 8) return p.out_dtype == DT_F32 ? launch_mode3<MODE, true, 8>(u, p, smem_bytes, st) : launch_mode3<MODE, false, 8>(u, p, smem_bytes, st);
    if (p.dual == 4) return p.out_dtype == DT_F32 ? launch_mode3<MODE, true, 4>(u, p, smem_bytes, st) : launch_mode3<MODE, false, 4>(u, p, smem_bytes, st);
    if (p.dual == 3) return p.out_dtype == DT_F32 ? launch_mode3<MODE, true, 3>(u, p, smem_bytes, st) : launch_mode3<MODE, false, 3>(u, p, smem_bytes, st);
    if (p.dual == 1) return p.out_dtype == DT_F32 ? launch_mode3<MODE, true, 1>(u, p, smem_bytes, st) : launch_mode3<MODE, false, 1>(u, p, smem_bytes, st);
    if (p.dual == 2) return p.out_dtype == DT_F32 ? launch_mode3<MODE, true, 2>(u, p, smem_bytes, st) : launch_mode3<MODE, false, 2>(u, p, smem_bytes, st);
  }
  return p.out_dtype == DT_F32 ? launch_mode3<MODE, true, 0>(u, p, smem_bytes, st) : launch_mode3<MODE, false, 0>(u, p, smem_bytes, st);
}

// One launch over the M tiles [mt_begin, mt_begin + mt_count) of the layer (mt_count <= 0: all of them).
static int launch_range(const UmmaConv& u, const ConvDesc& d, cudaStream_t st, int mt_begin, int mt_count, bool allow_wide) {
  if (!u.enabled) return fail(YOLO_E_STATE, "umma: tensor maps not built");
  if (d.N > u.max_batch) return fail(YOLO_E_SHAPE, "umma: batch %d exceeds the tensor map's %d", d.N, u.max_batch);
  if (g_num_sms == 0) {
    int dev = 0;
    YB_CUDA(cudaGetDevice(&dev));
    YB_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int np = planes_of(u.precision);
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.M = d.N * d.Ho * d.Wo;
  p.Cout = d.Cout;
  p.BN = u.bn_tile;
  p.bk = u.bk;
  p.n_tiles_n = (d.Cout + p.BN - 1) / p.BN;
  p.n_tiles = ((p.M + TILE_M - 1) / TILE_M) * p.n_tiles_n;
  p.taps = d.kh * d.kw; p.kw = d.kw; p.cin_blocks = u.c32i ? 1 : d.Cin / p.bk;
  p.Ho = d.Ho; p.Wo = d.Wo; p.stride = d.stride; p.pad = d.pad;
  p.in_coff = d.in_coff;
  p.a_plane_n = u.max_batch;
  p.a_tiled = u.a_tiled ? 1 : 0;
  p.a_plane_rows = u.a_plane_rows;
  p.b_plane_rows = p.n_tiles_n * p.BN;
  // dual-M tiles (fp16x3 only, EXPERIMENTAL, off by default; YOLO_B200_DUAL=1 enables): sharing the weight tile between two
  // M tiles cuts the TMA bytes per MMA by 25 % but measured no gain (head 3x3: 753 vs 760 us) - the main loop is bound
  // by SHARED-MEMORY bandwidth (MMA operand reads + TMA fill = 213 B/clk at 128x128 vs the 128 B/clk port), which this
  // does not change enough; see profiles/r1_ncu_summary.md.  The real fix is cta_group::2 with 256x256 tiles.
  const int m_tiles = mt_count > 0 ? mt_count : (p.M + TILE_M - 1) / TILE_M;
  p.mt_begin = mt_begin;
  p.n_tiles = m_tiles * p.n_tiles_n;
  p.dual = 0;
  if (const char* de = getenv("YOLO_B200_DUAL")) p.dual = (de[0] == '1' && mode_of(u.precision) == 2 && m_tiles >= 2) ? 1 : 0;
  // 2-CTA pairs (cta_group::2, 256 x BN): fp16x3, needs the half-height weight map and at least one full pair of M tiles
  const char* pe = getenv("YOLO_B200_PAIR");
  if (!p.dual && !u.a_tiled && mode_of(u.precision) == 2 && u.has_map_b2 && m_tiles >= 2 && p.BN % 32 == 0 && pe) {
    if (pe[0] == '1') p.dual = 2;
    if (pe[0] == '2' && p.n_tiles_n % 2 == 0) p.dual = 3;               // 256 x 2BN: pairs of N tiles share the activation tiles
  }
  // multicast clusters (KIND 4): two CTAs with neighbouring M tiles fetch half of the weight tile each and multicast it
  const char* me = getenv("YOLO_B200_MCAST");
  if (!p.dual && mode_of(u.precision) == 2 && u.has_map_b2 && m_tiles >= 2 && p.BN % 32 == 0 && me && me[0] == '1') p.dual = 4;
  // wide tiles (KIND 5): 128 x 256, merged accumulation
  // (default where Cout % 256 == 0: shared-memory bytes per flop drop by 25 % - the main loop is bound by the shared-memory
  //  port, MMA operand reads + TMA fill - measured 1.36x on the Darknet-53 step; YOLO_B200_WIDE=0 switches it off)
  const char* we = getenv("YOLO_B200_WIDE");
  if (!p.dual && allow_wide && u.has_map_bw && d.out_dtype != DT_F32 && !(we && we[0] == '0')) {
    p.dual = 5;
    p.BN = 256;
    p.n_tiles_n = d.Cout / 256;
    p.n_tiles = m_tiles * p.n_tiles_n;
  }
  if (u.c32i) p.dual = 7;
  // wide pairs (KIND 6, default where at least two M tiles exist; YOLO_B200_PAIRWIDE=0 keeps the single-CTA wide tiles):
  // 256 x 256 per CTA pair, each CTA stages half of the weight rows -> a third less fill per SM and three pipeline stages.
  // Pays off only together with two-k-block partials in hh-last order (below): every partial hand-off crosses the cluster.
  const char* pw = getenv("YOLO_B200_PAIRWIDE");
  if (p.dual == 5 && mt_count <= 0 && m_tiles >= 2 && u.bn_tile == 128 && !(pw && pw[0] == '0')) p.dual = 6;
  if (p.dual == 3) p.n_tiles_n /= 2;                                    // scheduling units per row of the tile grid
  // EXPERIMENT (YOLO_B200_NARROWMERGE=1): merged accumulation + hh-last partials on the 128-column tiles as well
  if (const char* nm = getenv("YOLO_B200_NARROWMERGE")) { if (nm[0] == '1' && p.dual == 0 && mode_of(u.precision) == 2 && p.bk == 64) p.dual = 8; }
  if (p.dual && p.dual != 5 && p.dual != 7 && p.dual != 8) p.n_tiles = ((m_tiles + 1) / 2) * p.n_tiles_n;
  const int stage_bytes = p.dual == 7 ? TILE_M * p.bk * 2 + np * p.BN * p.bk * 2 : np * ((p.dual == 1 ? 2 : 1) * TILE_M * p.bk * 2 + (p.dual == 3 ? 2 : 1) * ((p.dual == 2 || p.dual == 3 || p.dual == 6) ? p.BN / 2 : p.BN) * p.bk * 2);
  const int aux_bytes = 16 * MAX_STAGES + 128 + 2 * 2 * p.BN * 4 + 64;
  // 8 MMAs of the leading product per TMEM partial.  Longer partials are ~3 % faster but their truncation bias is
  // systematic (same sign on every output) and compounds through the layers: flush 8 -> Darknet-53 head error 6.6e-4
  // instead of 2.9e-4 (profiles/r1_parity_report.txt).  Env YOLO_B200_FLUSH overrides for experiments.
  p.flush = p.bk == 64 ? 2 : 4;
  // merged accumulation (default): every plane pair accumulates in the partial buffer, correction products first, one
  // k-block per partial.  Measured better AND faster than the separate correction accumulator (r1_ncu_summary.md).
  if (p.dual == 5 || p.dual == 6) p.flush = 1;
  if (p.dual == 7) p.flush = 4;                                          // two hi*hi MMAs per tap -> 8 per partial                                          // merged accumulation: one k-block (12 MMAs) per partial
  if (const char* de2 = getenv("YOLO_B200_DBG_PAIRS")) p.dbg_pairs = atoi(de2);
  if (const char* ns = getenv("YOLO_B200_DBG_NOSTORE")) p.dbg_nostore = atoi(ns);
  // hh-last partials (default on the pair kind): one partial = two k-blocks, the correction products of both first, then the
  // leading ones.  Measured (3x3 512->1024 @26^2): KIND 6 587 -> 503 us, KIND 5 589 -> 590 us (it has only two stages to hold);
  // Darknet-53 head error 2.4e-4 -> 2.7e-4, step 14.9 -> 13.8 ms.  YOLO_B200_HHLAST=0|1 overrides.
  {
    const char* hl = getenv("YOLO_B200_HHLAST");
    const bool on = hl ? hl[0] == '1' : (p.dual == 6 || p.dual == 8);
    if (on && (p.dual == 5 || p.dual == 6 || p.dual == 8)) { p.hh_last = 1; p.flush = 2; }
  }
  if (const char* bs = getenv("YOLO_B200_BSPLIT")) p.b_split = (bs[0] == '1' && p.dual == 5 && u.bn_tile == 128) ? 1 : 0;
  if (const char* pf = getenv("YOLO_B200_PREFETCH")) p.prefetch = atoi(pf);
  if (const char* fe = getenv("YOLO_B200_FLUSH")) { int f = atoi(fe); if (f >= 1 && f <= 64) p.flush = f; }
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
  if (d.out_dtype != DT_F32 && (((d.out_cpitch | d.out_coff) & 7) || d.Cout % 8)) return fail(YOLO_E_UNSUPPORTED, "umma: 16-bit output needs 16-byte aligned channel slices and Cout %% 8 == 0");
  if (d.res && ((d.res_cpitch | d.res_coff) & 7)) return fail(YOLO_E_UNSUPPORTED, "umma: residual needs 16-byte aligned channel slices");
  if (d.res && (d.Cout % 32 || d.out_dtype == DT_F32)) return fail(YOLO_E_UNSUPPORTED, "umma: residual needs Cout %% 32 == 0 and a 16-bit activation format");
  const int smem_bytes = 1024 + stages * stage_bytes + aux_bytes;
  switch (mode_of(u.precision)) {
    case 0: return launch_mode<0>(u, p, smem_bytes, st);
    case 1: return launch_mode<1>(u, p, smem_bytes, st);
    default: return launch_mode<2>(u, p, smem_bytes, st);
  }
}

// Wave quantisation (EXPERIMENT, YOLO_B200_SPLIT=1; off by default).  A persistent launch runs ceil(tiles / SMs) rounds of
// tiles, and with 128 x 256 tiles the layers of the 13^2 and 26^2 maps have only 1.2 - 4.6 rounds: the last, partly filled
// round costs as much as a full one.  A layer can be split along M: as many FULL rounds of wide tiles as fit, and the remaining
// M tiles as 128 x 128 tiles in a second launch.  Measured: a narrow tile costs 0.75-0.85 of a wide one (not 0.5), and the
// second launch ~10 us, so the step got 3 % SLOWER with the split (15.45 vs 14.94 ms); the planner is kept for the record.
int launch_conv_umma(const UmmaConv& u, const ConvDesc& d, cudaStream_t st) {
  const char* we = getenv("YOLO_B200_WIDE");
  const char* se = getenv("YOLO_B200_SPLIT");
  const bool wide_ok = u.enabled && u.has_map_bw && d.out_dtype != DT_F32 && !(we && we[0] == '0') && !getenv("YOLO_B200_DUAL") &&
                       !getenv("YOLO_B200_PAIR") && !getenv("YOLO_B200_MCAST");
  if (!wide_ok || !(se && se[0] == '1')) return launch_range(u, d, st, 0, 0, true);
  if (g_num_sms == 0) {
    int dev = 0;
    YB_CUDA(cudaGetDevice(&dev));
    YB_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int m_tiles = (d.N * d.Ho * d.Wo + TILE_M - 1) / TILE_M;
  const int n_w = d.Cout / 256, n_n = 2 * n_w, sms = g_num_sms;
  const double c_n = 0.78;                                            // narrow tile time / wide tile time (measured, head 3x3 layers)
  const double second_launch = 10.0 / (1.6 * d.kh * d.kw * (d.Cin / 64)); // ~10 us of launch + prologue + tail, in wide-tile times
  auto rounds = [&](int tiles) { return (tiles + sms - 1) / sms; };
  int best_mw = m_tiles;
  double best = rounds(m_tiles * n_w) * 1.0;
  for (int w = 0; w <= rounds(m_tiles * n_w); ++w) {
    int mw = (int)(((long long)w * sms) / n_w);
    if (mw > m_tiles) mw = m_tiles;
    const double cost = rounds(mw * n_w) * 1.0 + rounds((m_tiles - mw) * n_n) * c_n + ((mw > 0 && mw < m_tiles) ? second_launch : 0.0);
    if (cost < best - 1e-9) { best = cost; best_mw = mw; }
  }
  if (best_mw == m_tiles) return launch_range(u, d, st, 0, 0, true);
  if (best_mw == 0) return launch_range(u, d, st, 0, 0, false);
  int rc = launch_range(u, d, st, 0, best_mw, true);
  if (rc) return rc;
  return launch_range(u, d, st, best_mw, m_tiles - best_mw, false);
}

}  // namespace yb
