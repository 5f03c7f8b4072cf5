// Fused anchor decode + box selection for sm_100a.
//
// Replaces (reference, file:line): merge_and_slice car/YOLO.py:841-849, _init_syxhw :123-155 (tables are
// recomputed from the flat index instead of being read), _yxhw_to_ltrb :552-566 and the per-image python
// loop of predict :568-597 -- about 30 tiny MXNet kernels plus a device->host sync per image -- by ONE
// launch.  One thread-block cluster of 8 CTAs per image: every CTA scans 1/8 of the image's boxes
// (objectness channel only), the partial results meet in CTA 0's shared memory through DSMEM, CTA 0
// finishes (top-1 gather, or sort + class-aware greedy NMS with warp-ballot suppression).
//
// Bit-exact selection: sigmoid is evaluated as the oracle defines it (correctly rounded expf, IEEE fp32
// add and divide, no FMA contraction), ties resolve to the lowest flat index like MXNet's argmax.
#include <cooperative_groups.h>
#include <limits.h>
#include <math_constants.h>

#include "common.cuh"
#include "decode_geom.cuh"

namespace cg = cooperative_groups;

namespace yb {

constexpr int kCluster = 8;        // CTAs per image
constexpr int kThreads = 256;
constexpr int kMaxRaw = 4096;      // raw candidates gathered per image (NMS mode)
constexpr int kMaxCand = 1024;     // candidates entering suppression

__device__ __forceinline__ float exp_cr(float x) {      // correctly rounded fp32 exp (oracle: exp32)
  return (float)exp((double)x);
}
__device__ __forceinline__ float sigmoid_exact(float x) {   // oracle: sigmoid32
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, exp_cr(-x)));
}

__device__ __forceinline__ const float* row_ptr(const DecodeDev& g, int b, int j, int& s, int& local) {
  s = 0;
#pragma unroll
  for (int k = 1; k < YOLO_MAX_SCALES; ++k)
    if (k < g.n_scales && j >= g.off[k]) s = k;
  local = j - g.off[s];
  return g.head[s] + ((size_t)b * g.boxes[s] + local) * g.C;
}

// ltrb of flat box j, arithmetic in the reference's operation order (car/YOLO.py:552-566).
__device__ __forceinline__ float4 box_ltrb(const DecodeDev& g, const float* row, int s, int local) {
  int cell = local / g.A, a = local - cell * g.A;
  int cy = cell / g.ws[s], cx = cell - cy * g.ws[s];
  float st = g.step[s];
  float y0 = (float)cy * st, x0 = (float)cx * st;
  float by = __fdiv_rn(__fadd_rn(__fmul_rn(sigmoid_exact(row[1]), st), y0), g.img_h);
  float bx = __fdiv_rn(__fadd_rn(__fmul_rn(sigmoid_exact(row[2]), st), x0), g.img_w);
  float bh = __fmul_rn(exp_cr(row[3]), g.anc[s][a][0]);
  float bw = __fmul_rn(exp_cr(row[4]), g.anc[s][a][1]);
  float bh2 = __fmul_rn(bh, 0.5f), bw2 = __fmul_rn(bw, 0.5f);
  return make_float4(__fsub_rn(bx, bw2), __fsub_rn(by, bh2), __fadd_rn(bx, bw2), __fadd_rn(by, bh2));
}

// yolo_gluon.get_iou (yolo_modules/yolo_gluon.py:158-167) in ltrb form, fp32 without contraction.
__device__ __forceinline__ float iou_ltrb(float4 a, float4 b) {
  float il = fmaxf(a.x, b.x), it = fmaxf(a.y, b.y), ir = fminf(a.z, b.z), ib = fminf(a.w, b.w);
  float iw = fmaxf(__fsub_rn(ir, il), 0.f), ih = fmaxf(__fsub_rn(ib, it), 0.f);
  float inter = __fmul_rn(iw, ih);
  float aa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  float ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
}

// iou_ltrb(a, b) > thr, bit-identical to the division form: the product form decides whenever it is not within rounding distance of the
// threshold (almost always); only then is the division evaluated.
__device__ __forceinline__ bool iou_above(float4 a, float4 b, float thr) {
  float il = fmaxf(a.x, b.x), it = fmaxf(a.y, b.y), ir = fminf(a.z, b.z), ib = fminf(a.w, b.w);
  float iw = fmaxf(__fsub_rn(ir, il), 0.f), ih = fmaxf(__fsub_rn(ib, it), 0.f);
  float inter = __fmul_rn(iw, ih);
  if (inter <= 0.f && thr >= 0.f) return false;
  float aa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  float ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  float uni = __fsub_rn(__fadd_rn(aa, ab), inter);
  float d = inter - thr * uni;
  if (uni > 0.f && fabsf(d) > 4e-6f * fabsf(uni)) return d > 0.f;
  return __fdiv_rn(inter, uni) > thr;
}

__device__ __forceinline__ void better(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// Writes one output row in predict()'s format from the gathered head row.
__device__ __forceinline__ void write_row(const DecodeDev& g, const float* row, float score, float4 box,
                                          float* out, int tid, int nthreads) {
  for (int c = tid; c < g.C; c += nthreads) {
    float v;
    if (c == 0) v = score;
    else if (c == 1) v = __fmul_rn(__fadd_rn(box.y, box.w), 0.5f);   // y = (t+b)/2
    else if (c == 2) v = __fmul_rn(__fadd_rn(box.x, box.z), 0.5f);   // x = (l+r)/2
    else if (c == 3) v = __fsub_rn(box.w, box.y);                    // h = b-t
    else if (c == 4) v = __fsub_rn(box.z, box.x);                    // w = r-l
    else v = row[c];
    out[c] = v;
  }
}

struct NmsDev {
  float score_thr, iou_thr;
  int max_out, max_cand;
};

// suppression bitmask, upper triangle only: row i keeps the words w >= i/32 (bits j > i), packed row after row
constexpr int kMaskWords = kMaxCand * (kMaxCand / 32) - 32 * ((kMaxCand / 32) * (kMaxCand / 32 - 1) / 2);      // 16896 for 1024 candidates
__device__ __forceinline__ int mask_row_offset(int i, int nw) {
  const int g = i >> 5;
  return i * nw - (32 * (g * (g - 1) / 2) + (i - 32 * g) * g);
}
struct NmsSmem {
  unsigned long long sorted[kMaxCand];   // the best K candidates in (score desc, index asc) order
  float4 box[kMaxCand];
  unsigned short cls[kMaxCand];
  unsigned short keep[kMaxCand];
  union {                                // the raw candidate list is dead once the rank sort has run: the bitmask reuses its space
    unsigned long long keys[kMaxRaw];    // raw candidates (rank 0 gathers them; every CTA takes a copy)
    unsigned mask[kMaskWords];           // rank 0: bit j of row i = "i suppresses j" (j > i, same class, IoU > thr)
  };
};                                       // 96 KB: two CTAs per SM, so the 8 x batch CTAs of 32 images run in one wave

template <int MODE>   // 0 = top-1 (reference semantics), 1 = NMS extension
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads)
decode_kernel(const __grid_constant__ DecodeDev g, const NmsDev np, float* __restrict__ out_rows,
              int* __restrict__ out_idx, int* __restrict__ out_count) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  __shared__ float s_wv[kThreads / 32];
  __shared__ int s_wi[kThreads / 32];
  __shared__ float s_cv[kCluster];      // valid in rank 0: per-CTA partial maxima
  __shared__ int s_ci[kCluster];
  __shared__ int s_count;               // valid in rank 0: raw candidate count (NMS)
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  NmsSmem* ns = reinterpret_cast<NmsSmem*>(dyn_smem);

  if (MODE == 1) {
    if (tid == 0) s_count = 0;
    cluster.sync();                     // rank 0's counter is zero before anybody appends
  }
  int* r0_count = MODE == 1 ? cluster.map_shared_rank(&s_count, 0) : nullptr;
  unsigned long long* r0_keys = MODE == 1 ? cluster.map_shared_rank(ns->keys, 0) : nullptr;

  // ---- scan: objectness of this CTA's slice of the image ------------------------------------
  const int per = (g.total + kCluster - 1) / kCluster;
  const int j0 = rank * per, j1 = min(g.total, j0 + per);
  float best = -1.f;
  int bidx = INT_MAX;
  for (int jb = j0; jb < j1; jb += kThreads) {            // uniform trip count: the warp stays converged for the ballot below
    const int j = jb + tid;
    float sc = -1.f;
    if (j < j1) {
      int s, local;
      const float* row = row_ptr(g, b, j, s, local);
      sc = sigmoid_exact(__ldg(row));
      if (sc > best) { best = sc; bidx = j; }
    }
    if (MODE == 1) {
      // candidates above the threshold are appended to rank 0's list: ONE remote atomic per warp and iteration (warp-aggregated),
      // not one per candidate - the serialised DSMEM atomics were the longest phase at ~1000 candidates per image
      const bool cand = j < j1 && sc > np.score_thr;
      const unsigned m = __ballot_sync(0xffffffffu, cand);
      if (m) {
        const int leader = __ffs(m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(r0_count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        const int slot = base + __popc(m & ((1u << lane) - 1u));
        if (cand && slot < kMaxRaw)
          r0_keys[slot] = ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(sc)) << 32) | (unsigned)j;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    better(best, bidx, ov, oi);
  }
  if (lane == 0) { s_wv[warp] = best; s_wi[warp] = bidx; }
  __syncthreads();
  if (warp == 0) {
    best = lane < kThreads / 32 ? s_wv[lane] : -1.f;
    bidx = lane < kThreads / 32 ? s_wi[lane] : INT_MAX;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      better(best, bidx, ov, oi);
    }
    if (lane == 0) {
      *cluster.map_shared_rank(&s_cv[rank], 0) = best;
      *cluster.map_shared_rank(&s_ci[rank], 0) = bidx;
    }
  }
  cluster.sync();                       // partials (and candidates) have landed in rank 0
  if (MODE == 0 && rank != 0) return;

  // ---- final top-1 (rank 0 holds the per-CTA partial maxima; in NMS mode it only seeds an empty candidate list) ----------------------
  if (rank == 0) {
    best = lane < kCluster ? s_cv[lane] : -1.f;
    bidx = lane < kCluster ? s_ci[lane] : INT_MAX;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      better(best, bidx, ov, oi);
    }
    best = __shfl_sync(0xffffffffu, best, 0);
    bidx = __shfl_sync(0xffffffffu, bidx, 0);
    if (bidx == INT_MAX) bidx = 0;        // all-NaN objectness: defined as index 0
  }

  if (MODE == 0) {
    int s, local;
    const float* row = row_ptr(g, b, bidx, s, local);
    float4 box = box_ltrb(g, row, s, local);
    write_row(g, row, sigmoid_exact(row[0]), box, out_rows + (size_t)b * g.C, tid, kThreads);
    if (tid == 0) out_idx[b] = bidx;
    return;
  }

  // ---- NMS: the whole cluster works on the image ------------------------------------------------------------------------------------
  // 1. every CTA copies the raw candidates; 2. rank sort: the rank of a candidate = number of smaller keys (keys are unique), computed
  // for a stride-8 slice per CTA and scattered into rank 0's `sorted`; 3. every CTA decodes the K best boxes; 4. all-pairs suppression
  // bitmask, rows interleaved over the CTAs, written into rank 0's shared memory; 5. one warp of rank 0 runs the greedy scan over the
  // bitmask rows of the KEPT boxes only (a suppressed box suppresses nothing) - identical to the sequential definition, without a
  // block barrier per kept box.
  int raw = *r0_count;                                                 // DSMEM read of rank 0's counter
  if (raw > kMaxRaw) {                  // cannot order more than kMaxRaw candidates exactly
    if (rank == 0 && tid == 0) out_count[b] = -raw;
    cluster.sync();                     // nobody leaves while a peer may still read its shared memory
    return;
  }
  if (raw == 0) {                       // nothing above the threshold: keep the top-1 alone
    if (rank == 0 && tid == 0)
      ns->keys[0] = ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(best)) << 32) | (unsigned)bidx;
    raw = 1;
    cluster.sync();
  }
  const int K = min(raw, min(np.max_cand, kMaxCand));
  if (rank != 0)
    for (int i = tid; i < raw; i += kThreads) ns->keys[i] = r0_keys[i];
  __syncthreads();
  {
    unsigned long long* r0_sorted = cluster.map_shared_rank(ns->sorted, 0);
    for (int i = rank + kCluster * tid; i < raw; i += kCluster * kThreads) {
      const unsigned long long key = ns->keys[i];
      int r = 0;
      for (int j = 0; j < raw; ++j) r += ns->keys[j] < key;             // broadcast shared-memory reads
      if (r < K) r0_sorted[r] = key;
    }
  }
  cluster.sync();                       // rank 0's `sorted` is complete
  {
    const unsigned long long* r0_sorted = cluster.map_shared_rank(ns->sorted, 0);
    for (int i = tid; i < K; i += kThreads) {
      const unsigned long long key = r0_sorted[i];
      if (rank != 0) ns->sorted[i] = key;
      const int j = (int)(key & 0xFFFFFFFFu);
      int s, local;
      const float* row = row_ptr(g, b, j, s, local);
      ns->box[i] = box_ltrb(g, row, s, local);
      // np.argmax over the class logits (first maximum); the row is read with independent 8-byte loads first (rows are 8-byte aligned
      // when C is even) so that their latencies overlap instead of one dependent load per compare
      int c = 0;
      float cv = -CUDART_INF_F;
      if ((g.C & 1) == 0 && g.C <= 96) {
        float2 r2[48];
#pragma unroll
        for (int q = 0; q < 48; ++q)
          if (2 * q < g.C) r2[q] = __ldg(reinterpret_cast<const float2*>(row) + q);
#pragma unroll
        for (int q = 3; q < 48; ++q) {
          if (2 * q < g.C) {
            if (r2[q].x > cv) { cv = r2[q].x; c = 2 * q - 6; }
            if (r2[q].y > cv) { cv = r2[q].y; c = 2 * q + 1 - 6; }
          }
        }
      } else {
        for (int q = 6; q < g.C; ++q) {
          float v = row[q];
          if (v > cv) { cv = v; c = q - 6; }
        }
      }
      ns->cls[i] = (unsigned short)c;
    }
  }
  __syncthreads();
  const int nw = (K + 31) >> 5;                                         // bitmask words per row
  {
    unsigned* r0_mask = cluster.map_shared_rank(ns->mask, 0);
    // (row, word) pairs of this CTA's rows i = rank, rank + 8, ...; the threads of a warp share the WORD and differ in the row, so the
    // candidate-side loads (cls[j], box[j]) are shared-memory broadcasts instead of 32-way bank conflicts
    const int R = (K - rank + kCluster - 1) / kCluster;
    const int pairs = R > 0 ? R * nw : 0;
    for (int e = tid; e < pairs; e += kThreads) {
      const int w = e / R, i = (e - w * R) * kCluster + rank;
      if (w < (i >> 5)) continue;                                       // lower triangle: not stored
      unsigned m = 0;
      const float4 bi = ns->box[i];
      const int ci = ns->cls[i];
#pragma unroll 4
      for (int t = 0; t < 32; ++t) {
        const int j = w * 32 + t;
        if (j > i && j < K && ns->cls[j] == ci && iou_above(bi, ns->box[j], np.iou_thr)) m |= 1u << t;
      }
      r0_mask[mask_row_offset(i, nw) + w - (i >> 5)] = m;
    }
  }
  cluster.sync();                       // the bitmask has landed in rank 0; peers are done with remote memory
  if (rank != 0) return;
  __shared__ int s_nk;
  if (warp == 0) {
    unsigned removed = lane < nw ? 0u : 0xFFFFFFFFu;   // lane l holds word l of the "suppressed" bitmap (bits >= K count as suppressed)
    if (lane == nw - 1 && (K & 31)) removed |= ~((1u << (K & 31)) - 1u);
    int nk = 0;
    int w = 0;                          // current word; candidates before it are decided
    unsigned done = 0;                  // bits of word w already visited
    while (w < nw) {
      const unsigned alive = ~(__shfl_sync(0xffffffffu, removed, w) | done);
      if (!alive) { ++w; done = 0; continue; }                         // warp-uniform: jump over suppressed runs 32 at a time
      const int bit = __ffs(alive) - 1, i = w * 32 + bit;
      done |= (bit == 31) ? 0xFFFFFFFFu : ((2u << bit) - 1u);
      if (lane == 0) ns->keep[nk] = i;
      ++nk;
      if (nk >= np.max_out) break;
      if (lane >= (i >> 5) && lane < nw) removed |= ns->mask[mask_row_offset(i, nw) + lane - (i >> 5)];
    }
    if (lane == 0) s_nk = nk;
  }
  __syncthreads();
  const int nk = s_nk;
  for (int e = tid; e < nk * g.C; e += kThreads) {                     // (kept box, channel) pairs: independent loads in flight
    const int k = e / g.C, c = e - k * g.C;
    const int i = ns->keep[k];
    const unsigned long long key = ns->sorted[i];
    const int j = (int)(key & 0xFFFFFFFFu);
    const float4 box = ns->box[i];
    float v;
    if (c == 0) v = __uint_as_float(0xFFFFFFFFu - (unsigned)(key >> 32));
    else if (c == 1) v = __fmul_rn(__fadd_rn(box.y, box.w), 0.5f);
    else if (c == 2) v = __fmul_rn(__fadd_rn(box.x, box.z), 0.5f);
    else if (c == 3) v = __fsub_rn(box.w, box.y);
    else if (c == 4) v = __fsub_rn(box.z, box.x);
    else { int s, local; v = row_ptr(g, b, j, s, local)[c]; }
    out_rows[((size_t)b * np.max_out + k) * g.C + c] = v;
    if (c == 0) out_idx[(size_t)b * np.max_out + k] = j;
  }
  if (tid == 0) out_count[b] = nk;
}

// ---- licence-plate pose decode ---------------------------------------------------------------
// mode 0: car_and_LP/YOLO.py:133-169 (NHWC map, argmax of sigmoid(score), 7 outputs)
// mode 1: licence_plate/LP_detection.py:147-162 (NCHW map, argmax of the raw score, ch outputs)
__global__ void __launch_bounds__(kThreads)
decode_lp_kernel(const float* __restrict__ lp, int n, int ch, int mode, float r0, float r1, float r2,
                 float* __restrict__ out_rows, int* __restrict__ out_idx) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* base = lp + (size_t)b * n * ch;
  const size_t cell_stride = mode == 0 ? ch : 1, ch_stride = mode == 0 ? 1 : n;
  __shared__ float s_wv[kThreads / 32];
  __shared__ int s_wi[kThreads / 32];
  float best = -CUDART_INF_F;
  int bidx = INT_MAX;
  for (int j = tid; j < n; j += kThreads) {
    float v = __ldg(base + j * cell_stride);
    if (mode == 0) v = sigmoid_exact(v);
    if (v > best) { best = v; bidx = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    better(best, bidx, ov, oi);
  }
  if (lane == 0) { s_wv[warp] = best; s_wi[warp] = bidx; }
  __syncthreads();
  if (warp != 0) return;
  best = lane < kThreads / 32 ? s_wv[lane] : -CUDART_INF_F;
  bidx = lane < kThreads / 32 ? s_wi[lane] : INT_MAX;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    better(best, bidx, ov, oi);
  }
  bidx = __shfl_sync(0xffffffffu, bidx, 0);
  if (bidx == INT_MAX) bidx = 0;
  const int nout = mode == 0 ? 7 : ch;
  if (lane < nout) {
    float v = base[bidx * cell_stride + lane * ch_stride];
    float o;
    if (lane == 0) o = sigmoid_exact(v);
    else if (lane < 4) o = __fmul_rn(v, 1000.f);
    else if (lane < 7) {
      float rm = lane == 4 ? r0 : (lane == 5 ? r1 : r2);
      float d = __fmul_rn(__fmul_rn(__fsub_rn(sigmoid_exact(v), 0.5f), 2.f), rm);
      o = __fdiv_rn(__fmul_rn(d, 3.14159274101257324f), 180.f);
    } else o = v;
    out_rows[(size_t)b * nout + lane] = o;
  }
  if (lane == 0 && out_idx) out_idx[b] = bidx;
}

}  // namespace yb

using namespace yb;

extern "C" int yolo_decode_top1(const yolo_decode_geom* g, const void* const* heads, int batch,
                                float* out_rows, int32_t* out_idx, void* stream) {
  DecodeDev d;
  int rc = make_dev(g, heads, d);
  if (rc) return rc;
  if (batch < 0 || !out_rows || !out_idx) return fail(YOLO_E_BADARG, "decode_top1: bad batch/outputs");
  if (batch == 0) return YOLO_OK;
  NmsDev np{0.f, 0.f, 0, 0};
  dim3 grid(kCluster, batch);
  decode_kernel<0><<<grid, kThreads, 0, (cudaStream_t)stream>>>(d, np, out_rows, out_idx, nullptr);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

extern "C" int yolo_decode_nms(const yolo_decode_geom* g, const void* const* heads, int batch,
                               const yolo_nms_params* p, float* out_rows, int32_t* out_idx, int32_t* out_count,
                               void* stream) {
  DecodeDev d;
  int rc = make_dev(g, heads, d);
  if (rc) return rc;
  if (!p || batch < 0 || !out_rows || !out_idx || !out_count) return fail(YOLO_E_BADARG, "decode_nms: bad arguments");
  if (p->max_out < 1 || p->max_cand < 1 || p->max_cand > kMaxCand || p->max_out > p->max_cand)
    return fail(YOLO_E_BADARG, "decode_nms: need 1 <= max_out <= max_cand <= %d", kMaxCand);
  if (batch == 0) return YOLO_OK;
  NmsDev np{p->score_thr, p->iou_thr, p->max_out, p->max_cand};
  rc = ensure_dyn_smem(reinterpret_cast<const void*>(&decode_kernel<1>), (int)sizeof(NmsSmem));
  if (rc) return rc;
  dim3 grid(kCluster, batch);
  decode_kernel<1><<<grid, kThreads, sizeof(NmsSmem), (cudaStream_t)stream>>>(d, np, out_rows, out_idx, out_count);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

extern "C" int yolo_decode_lp(const void* lp, int batch, int hs, int ws, int ch, int mode, const float r_max[3],
                              float* out_rows, int32_t* out_idx, void* stream) {
  if (!lp || !r_max || !out_rows || batch < 0 || hs < 1 || ws < 1) return fail(YOLO_E_BADARG, "decode_lp: bad arguments");
  if (ch < 7 || ch > 32 || (mode != 0 && mode != 1)) return fail(YOLO_E_BADARG, "decode_lp: ch=%d mode=%d unsupported", ch, mode);
  if (batch == 0) return YOLO_OK;
  decode_lp_kernel<<<batch, kThreads, 0, (cudaStream_t)stream>>>(static_cast<const float*>(lp), hs * ws, ch, mode,
                                                               r_max[0], r_max[1], r_max[2], out_rows, out_idx);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}
