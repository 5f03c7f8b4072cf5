// Training targets + losses + head gradients on the GPU.
//
// Replaces (reference, file:line):
//   _get_default_ltrb  car/YOLO.py:209-240   anchor boxes at the cell centres (recomputed from the flat index)
//   get_iou (mode 2)   yolo_modules/yolo_gluon.py:127-168
//   _find_best         car/YOLO.py:401-448   argmax IoU -> (pixel, anchor); the reference does this on the host with one
//                                            device->host sync PER LABEL (.asnumpy() at :404)
//   _loss_mask         car/YOLO.py:450-480   six dense (B, boxes, A, k) target tensors written element by element
//   _score_weight      car/YOLO.py:482-489
//   _get_loss          car/YOLO.py:491-498   LogisticLoss / HuberLoss x3 / SoftmaxCrossEntropyLoss, mean over non-batch axes
// by two launches that never materialise the dense targets: `assign_kernel` (one CTA per image) finds the matched box of
// every label and its regression targets, `loss_kernel` streams the head tensors once, accumulates the five per-image
// losses and (optionally) writes d(sum of losses)/d(head) - what `sum(losses).backward()` (car/YOLO.py:394) feeds into the net.
#include <limits.h>
#include <math_constants.h>

#include "common.cuh"
#include "decode_geom.cuh"

namespace yb {

constexpr int kLossThreads = 256;
constexpr int kMaxObj = 16;           // labels per image handled by the kernels

struct LossDev {
  float s_score, s_yx, s_hw, s_rot, s_cls;   // spec `scale` (rotate already 0 unless car_rotate)
  float w_pos, w_neg;
  int n_obj, n_class;
};

struct AssignRec {        // one per (image, label)
  int flat;               // matched flat box index, -1 = no label / overwritten by a later label on the same box
  float t[4];             // ty, tx, th, tw
};

__device__ __forceinline__ void better_first(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// anchor box k of _get_default_ltrb, fp32 op by op like the oracle
__device__ __forceinline__ float4 anchor_ltrb(const DecodeDev& g, int k, int& s_out, int& a_out) {
  int s = 0;
#pragma unroll
  for (int q = 1; q < YOLO_MAX_SCALES; ++q)
    if (q < g.n_scales && k >= g.off[q]) s = q;
  const int local = k - g.off[s];
  const int cell = local / g.A, a = local - cell * g.A;
  const int cy = cell / g.ws[s], cx = cell - cy * g.ws[s];
  const float sy = __fdiv_rn(g.step[s], g.img_h), sx = __fdiv_rn(g.step[s], g.img_w);
  const float yc = __fadd_rn(__fdiv_rn(sy, 2.f), __fmul_rn((float)cy, sy));
  const float xc = __fadd_rn(__fdiv_rn(sx, 2.f), __fmul_rn((float)cx, sx));
  const float hh = __fmul_rn(0.5f, g.anc[s][a][0]), hw = __fmul_rn(0.5f, g.anc[s][a][1]);
  s_out = s; a_out = a;
  return make_float4(__fsub_rn(xc, hw), __fsub_rn(yc, hh), __fadd_rn(xc, hw), __fadd_rn(yc, hh));
}

__global__ void __launch_bounds__(kLossThreads)
assign_kernel(const __grid_constant__ DecodeDev g, const LossDev lp, const float* __restrict__ labels, AssignRec* __restrict__ recs,
              int* __restrict__ out_assign) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lw = 6 + lp.n_class;
  __shared__ float s_wv[kLossThreads / 32];
  __shared__ int s_wi[kLossThreads / 32];
  __shared__ int s_flat[kMaxObj];
  for (int j = 0; j < lp.n_obj; ++j) {
    const float* L = labels + ((size_t)b * lp.n_obj + j) * lw;
    const float l0 = L[0];
    if (l0 < 0.f) {                                    // "if L[0] < 0: continue" (block-uniform)
      if (tid == 0) { s_flat[j] = -1; recs[(size_t)b * lp.n_obj + j].flat = -1; }
      continue;
    }
    const float Ly = L[1], Lx = L[2], Lh = L[3], Lw = L[4];
    const float l2 = __fsub_rn(Lx, __fdiv_rn(Lw, 2.f)), t2 = __fsub_rn(Ly, __fdiv_rn(Lh, 2.f));
    const float r2 = __fadd_rn(Lx, __fdiv_rn(Lw, 2.f)), b2 = __fadd_rn(Ly, __fdiv_rn(Lh, 2.f));
    const float ta = __fmul_rn(Lh, Lw);
    float best = -CUDART_INF_F;
    int bidx = INT_MAX;
    for (int k = tid; k < g.total; k += kLossThreads) {
      int s, a;
      const float4 q = anchor_ltrb(g, k, s, a);
      const float iw = fmaxf(__fsub_rn(fminf(r2, q.z), fmaxf(l2, q.x)), 0.f);
      const float ih = fmaxf(__fsub_rn(fminf(b2, q.w), fmaxf(t2, q.y)), 0.f);
      const float inter = __fmul_rn(iw, ih);
      const float pa = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
      const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(pa, ta), inter));
      if (iou > best) { best = iou; bidx = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      better_first(best, bidx, ov, oi);
    }
    if (lane == 0) { s_wv[warp] = best; s_wi[warp] = bidx; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kLossThreads / 32; ++w) better_first(best, bidx, s_wv[w], s_wi[w]);
      if (bidx == INT_MAX) bidx = 0;                   // all-NaN IoU: argmax returns 0
      int s, a;
      const float4 q = anchor_ltrb(g, bidx, s, a);
      AssignRec r;
      r.flat = bidx;
      const float st = g.step[s];
      const float by = __fsub_rn(Ly, __fdiv_rn(__fadd_rn(q.w, q.y), 2.f));
      float sy = __fadd_rn(__fdiv_rn(__fmul_rn(by, g.img_h), st), 0.5f);
      sy = fminf(fmaxf(sy, 0.0001f), 0.9999f);
      const float bx = __fsub_rn(Lx, __fdiv_rn(__fadd_rn(q.z, q.x), 2.f));
      float sx = __fadd_rn(__fdiv_rn(__fmul_rn(bx, g.img_w), st), 0.5f);
      sx = fminf(fmaxf(sx, 0.0001f), 0.9999f);
      r.t[0] = -logf(__fsub_rn(__fdiv_rn(1.f, sy), 1.f));   // nd_inv_sigmoid (yolo_gluon.py:365-367)
      r.t[1] = -logf(__fsub_rn(__fdiv_rn(1.f, sx), 1.f));
      r.t[2] = logf(__fdiv_rn(Lh, g.anc[s][a][0]));
      r.t[3] = logf(__fdiv_rn(Lw, g.anc[s][a][1]));
      recs[(size_t)b * lp.n_obj + j] = r;
      s_flat[j] = bidx;
    }
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) {
    // later labels overwrite earlier ones that matched the same box (dense-tensor assignment order, car/YOLO.py:466-478)
    for (int j = 0; j < lp.n_obj; ++j) {
      int f = s_flat[j];
      if (out_assign) out_assign[(size_t)b * lp.n_obj + j] = f;
      for (int j2 = j + 1; j2 < lp.n_obj && f >= 0; ++j2)
        if (s_flat[j2] == f) { recs[(size_t)b * lp.n_obj + j].flat = -1; break; }
    }
  }
}

__device__ __forceinline__ float softplus_neg_abs(float x) { return log1pf(expf(-fabsf(x))); }   // softrelu(-|x|)
__device__ __forceinline__ float huber(float d) { d = fabsf(d); return d > 1.f ? d - 0.5f : 0.5f * d * d; }
__device__ __forceinline__ float huber_grad(float d) { return fabsf(d) > 1.f ? (d > 0.f ? 1.f : -1.f) : d; }

// grid (chunks, B).  partial[b][chunk][5] = this block's sum of the five weighted loss terms.
__global__ void __launch_bounds__(kLossThreads)
loss_kernel(const __grid_constant__ DecodeDev g, const LossDev lp, const float* __restrict__ labels, const AssignRec* __restrict__ recs,
            float* __restrict__ partial, float* __restrict__ dh0, float* __restrict__ dh1, float* __restrict__ dh2) {
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lw = 6 + lp.n_class, C = g.C;
  __shared__ AssignRec s_rec[kMaxObj];
  __shared__ float s_red[kLossThreads / 32][5];
  if (tid < lp.n_obj) s_rec[tid] = recs[(size_t)b * lp.n_obj + tid];
  __syncthreads();
  const float N = (float)g.total;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  const int per = (g.total + gridDim.x - 1) / gridDim.x;
  const int k0 = blockIdx.x * per, k1 = min(g.total, k0 + per);
  for (int k = k0 + tid; k < k1; k += kLossThreads) {
    int s = 0;
#pragma unroll
    for (int q = 1; q < YOLO_MAX_SCALES; ++q)
      if (q < g.n_scales && k >= g.off[q]) s = q;
    const int local = k - g.off[s];
    const size_t roff = ((size_t)b * g.boxes[s] + local) * C;
    const float* row = g.head[s] + roff;
    float* drow = nullptr;
    if (dh0) drow = (s == 0 ? dh0 : (s == 1 ? dh1 : dh2)) + roff;
    int j = -1;
    for (int q = 0; q < lp.n_obj; ++q)
      if (s_rec[q].flat == k) j = q;
    // ---- objectness: LogisticLoss(binary) with weight where(mask, pos, neg) * scale['score'] ----
    const float p = __ldg(row);
    const float l = j >= 0 ? 1.f : -1.f;
    const float w = (j >= 0 ? lp.w_pos : lp.w_neg) * lp.s_score;
    const float pl = p * l;
    acc[0] += (fmaxf(-pl, 0.f) + softplus_neg_abs(pl)) * w;
    if (drow) {
      drow[0] = w * (-l) / (1.f + expf(pl)) / N;                  // d softplus(-pl)/dp = -l * sigmoid(-pl)
      if (j < 0)
        for (int c = 1; c < C; ++c) drow[c] = 0.f;
    }
    if (j >= 0) {
      const float* L = labels + ((size_t)b * lp.n_obj + j) * lw;
      const AssignRec& r = s_rec[j];
      // ---- Huber on (ty,tx), (th,tw), rotate; mask = 1 here ----
      float d0 = row[1] - r.t[0], d1 = row[2] - r.t[1], d2 = row[3] - r.t[2], d3 = row[4] - r.t[3], d4 = row[5] - L[5];
      acc[1] += (huber(d0) + huber(d1)) * lp.s_yx;
      acc[2] += (huber(d2) + huber(d3)) * lp.s_hw;
      acc[3] += huber(d4) * lp.s_rot;
      // ---- SoftmaxCrossEntropy(sparse_label=False): -sum(log_softmax(p) * label) ----
      float mx = -CUDART_INF_F;
      for (int c = 6; c < C; ++c) mx = fmaxf(mx, row[c]);
      float se = 0.f, sl = 0.f, dot = 0.f;
      for (int c = 6; c < C; ++c) { se += expf(row[c] - mx); sl += L[c]; dot += (row[c] - mx) * L[c]; }
      const float lse = logf(se);
      acc[4] += (lse * sl - dot) * lp.s_cls;
      if (drow) {
        drow[1] = lp.s_yx * huber_grad(d0) / (2.f * N);
        drow[2] = lp.s_yx * huber_grad(d1) / (2.f * N);
        drow[3] = lp.s_hw * huber_grad(d2) / (2.f * N);
        drow[4] = lp.s_hw * huber_grad(d3) / (2.f * N);
        drow[5] = lp.s_rot * huber_grad(d4) / N;
        for (int c = 6; c < C; ++c) drow[c] = lp.s_cls * (expf(row[c] - mx) / se * sl - L[c]) / N;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    float v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][q] = v;
  }
  __syncthreads();
  if (tid < 5) {
    float v = 0.f;
    for (int w = 0; w < kLossThreads / 32; ++w) v += s_red[w][tid];
    partial[((size_t)b * gridDim.x + blockIdx.x) * 5 + tid] = v;
  }
}

// losses[q][b] = mean over the non-batch axes (fixed summation order -> deterministic)
__global__ void loss_finalize_kernel(const float* __restrict__ partial, int chunks, int B, float total_boxes, float* __restrict__ losses) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 5 * B) return;
  const int q = i / B, b = i - q * B;
  float v = 0.f;
  for (int c = 0; c < chunks; ++c) v += partial[((size_t)b * chunks + c) * 5 + q];
  const float denom = (q == 1 || q == 2) ? 2.f * total_boxes : total_boxes;
  losses[(size_t)q * B + b] = v / denom;
}


// ---- licence-plate pose head: targets + losses + head gradient (LP_detection.py:259-313,354-360; car_and_LP/YOLO.py:124-131) ----------
// One CTA per image.  Labels: rows [flag, X, Y, Z (mm), r1, r2, r3 (rad), pixel x, pixel y, ..., class]; the target cell is
// (clip(int(L[8] / step)), clip(int(L[7] / step))) - `_find_best_LP`; a later label overwrites the pose of an earlier one in the same
// cell while the class one-hots accumulate (`_loss_mask_LP` never clears them).  Losses like the car head: LogisticLoss(binary) on the
// score with where(mask, pos, neg) weights, HuberLoss on xy / z / r, SoftmaxCrossEntropy on the class logits, each the mean over the
// non-batch axes.  The map is (B, Hs, Ws, ch) NHWC (car_and_LP) or (B, ch, Hs, Ws) NCHW (LPDenseNet output before `slice_out`).
struct LpLossDev {
  float s_score, s_xy, s_z, s_r, s_cls, w_pos, w_neg;
  int n_obj, n_lab, n_class, hs, ws, ch, nchw;
  float step, rmax[3];
};
struct LpRec { int cell; float t[6]; int cls; };

__global__ void __launch_bounds__(kLossThreads)
lp_loss_kernel(const float* __restrict__ lp, const float* __restrict__ labels, const LpLossDev p, float* __restrict__ losses, int B,
               float* __restrict__ dlp) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ LpRec s_rec[kMaxObj];
  __shared__ float s_red[kLossThreads / 32][5];
  if (tid < p.n_obj) {
    const float* L = labels + ((size_t)b * p.n_obj + tid) * p.n_lab;
    LpRec r;
    r.cell = -1; r.cls = -1;
    if (L[0] >= 0.f) {
      int hf = (int)__fdiv_rn(L[8], p.step), wf = (int)__fdiv_rn(L[7], p.step);          // int(): truncation toward zero
      hf = min(max(hf, 0), p.hs - 1); wf = min(max(wf, 0), p.ws - 1);
      r.cell = hf * p.ws + wf;
      r.t[0] = __fdiv_rn(L[1], 1000.f); r.t[1] = __fdiv_rn(L[2], 1000.f); r.t[2] = __fdiv_rn(L[3], 1000.f);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float rm = __fdiv_rn(__fmul_rn(p.rmax[i], 3.14159265358979323846f), 180.f);
        const float x = __fadd_rn(__fdiv_rn(__fdiv_rn(L[4 + i], rm), 2.f), 0.5f);
        r.t[3 + i] = -logf(__fsub_rn(__fdiv_rn(1.f, x), 1.f));                            // nd_inv_sigmoid (yolo_gluon.py:365-367)
      }
      const int c = (int)L[p.n_lab - 1];
      r.cls = (c >= 0 && c < p.n_class) ? c : -1;
    }
    s_rec[tid] = r;
  }
  __syncthreads();
  const int cells = p.hs * p.ws;
  const float N = (float)cells;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = tid; k < cells; k += kLossThreads) {
    int j = -1;
    unsigned clsmask = 0;
    for (int q = 0; q < p.n_obj; ++q)
      if (s_rec[q].cell == k) { j = q; if (s_rec[q].cls >= 0) clsmask |= 1u << s_rec[q].cls; }
    auto at = [&](int c) -> size_t { return p.nchw ? ((size_t)b * p.ch + c) * cells + k : ((size_t)b * cells + k) * p.ch + c; };
    float v[32];
    for (int c = 0; c < p.ch; ++c) v[c] = __ldg(lp + at(c));
    const float l = j >= 0 ? 1.f : -1.f;
    const float w = (j >= 0 ? p.w_pos : p.w_neg) * p.s_score;
    const float pl = v[0] * l;
    acc[0] += (fmaxf(-pl, 0.f) + softplus_neg_abs(pl)) * w;
    if (dlp) {
      dlp[at(0)] = w * (-l) / (1.f + expf(pl)) / N;
      if (j < 0)
        for (int c = 1; c < p.ch; ++c) dlp[at(c)] = 0.f;
    }
    if (j >= 0) {
      const LpRec& r = s_rec[j];
      const float d0 = v[1] - r.t[0], d1 = v[2] - r.t[1], d2 = v[3] - r.t[2], d3 = v[4] - r.t[3], d4 = v[5] - r.t[4], d5 = v[6] - r.t[5];
      acc[1] += (huber(d0) + huber(d1)) * p.s_xy;
      acc[2] += huber(d2) * p.s_z;
      acc[3] += (huber(d3) + huber(d4) + huber(d5)) * p.s_r;
      float mx = -CUDART_INF_F;
      for (int c = 7; c < p.ch; ++c) mx = fmaxf(mx, v[c]);
      float se = 0.f, sl = 0.f, dot = 0.f;
      for (int c = 7; c < p.ch; ++c) {
        const float lab = (clsmask >> (c - 7)) & 1u ? 1.f : 0.f;
        se += expf(v[c] - mx); sl += lab; dot += (v[c] - mx) * lab;
      }
      if (p.ch > 7) acc[4] += (logf(se) * sl - dot) * p.s_cls;
      if (dlp) {
        dlp[at(1)] = p.s_xy * huber_grad(d0) / (2.f * N);
        dlp[at(2)] = p.s_xy * huber_grad(d1) / (2.f * N);
        dlp[at(3)] = p.s_z * huber_grad(d2) / N;
        dlp[at(4)] = p.s_r * huber_grad(d3) / (3.f * N);
        dlp[at(5)] = p.s_r * huber_grad(d4) / (3.f * N);
        dlp[at(6)] = p.s_r * huber_grad(d5) / (3.f * N);
        for (int c = 7; c < p.ch; ++c) {
          const float lab = (clsmask >> (c - 7)) & 1u ? 1.f : 0.f;
          dlp[at(c)] = p.s_cls * (expf(v[c] - mx) / se * sl - lab) / N;
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    float x = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[warp][q] = x;
  }
  __syncthreads();
  if (tid < 5) {
    float x = 0.f;
    for (int w2 = 0; w2 < kLossThreads / 32; ++w2) x += s_red[w2][tid];
    const float denom = tid == 1 ? 2.f * N : (tid == 3 ? 3.f * N : N);
    losses[(size_t)tid * B + b] = x / denom;
  }
}

}  // namespace yb

using namespace yb;

extern "C" size_t yolo_loss_scratch_bytes(int batch, int n_obj) {
  return (size_t)batch * n_obj * sizeof(AssignRec) + (size_t)batch * 64 * 5 * sizeof(float) + 256;
}

extern "C" int yolo_loss_targets(const yolo_decode_geom* g, const void* const* heads, const float* labels, int batch, int n_obj,
                                 const yolo_loss_params* p, void* scratch, float* out_losses, void* const* dheads, int32_t* out_assign,
                                 void* stream) {
  DecodeDev d;
  int rc = make_dev(g, heads, d);
  if (rc) return rc;
  if (!labels || !p || !scratch || !out_losses || batch < 0 || n_obj < 1) return fail(YOLO_E_BADARG, "loss_targets: bad arguments");
  if (n_obj > kMaxObj) return fail(YOLO_E_UNSUPPORTED, "loss_targets: at most %d labels per image", kMaxObj);
  if (batch == 0) return YOLO_OK;
  LossDev lp;
  lp.s_score = p->scale_score; lp.s_yx = p->scale_box_yx; lp.s_hw = p->scale_box_hw;
  lp.s_rot = p->car_rotate ? p->scale_rotate : 0.f;           // _get_loss: rotate_lr = scale['rotate'] if car_rotate else 0
  lp.s_cls = p->scale_class;
  lp.w_pos = p->positive_weight; lp.w_neg = p->negative_weight;
  lp.n_obj = n_obj; lp.n_class = g->channels_per_anchor - 6;
  cudaStream_t st = (cudaStream_t)stream;
  AssignRec* recs = static_cast<AssignRec*>(scratch);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(scratch) + (((size_t)batch * n_obj * sizeof(AssignRec) + 255) & ~(size_t)255));
  const int chunks = 64;
  assign_kernel<<<batch, kLossThreads, 0, st>>>(d, lp, labels, recs, out_assign);
  float* dh[3] = {nullptr, nullptr, nullptr};
  if (dheads) {
    for (int s = 0; s < g->n_scales; ++s) {
      if (!dheads[s]) return fail(YOLO_E_BADARG, "loss_targets: dheads[%d] is null", s);
      dh[s] = static_cast<float*>(dheads[s]);
    }
  }
  loss_kernel<<<dim3(chunks, batch), kLossThreads, 0, st>>>(d, lp, labels, recs, partial, dh[0], dh[1], dh[2]);
  loss_finalize_kernel<<<(5 * batch + 127) / 128, 128, 0, st>>>(partial, chunks, batch, (float)d.total, out_losses);
  g_launches += 3;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

extern "C" int yolo_lp_loss_targets(const float* lp_map, int nchw, int batch, int hs, int ws, int ch, int step, const float r_max[3],
                                    const float* labels, int n_obj, int n_lab, const yolo_lp_loss_params* p, float* out_losses, float* dlp,
                                    void* stream) {
  if (!lp_map || !labels || !p || !out_losses || !r_max || batch < 0 || hs < 1 || ws < 1 || step < 1) return fail(YOLO_E_BADARG, "lp_loss_targets: bad arguments");
  if (ch < 7 || ch > 32 || n_lab < 10 || n_obj < 1) return fail(YOLO_E_BADARG, "lp_loss_targets: ch=%d n_lab=%d n_obj=%d unsupported", ch, n_lab, n_obj);
  if (n_obj > kMaxObj) return fail(YOLO_E_UNSUPPORTED, "lp_loss_targets: at most %d labels per image", kMaxObj);
  if (batch == 0) return YOLO_OK;
  LpLossDev d;
  d.s_score = p->scale_score; d.s_xy = p->scale_xy; d.s_z = p->scale_z; d.s_r = p->scale_r; d.s_cls = p->scale_class;
  d.w_pos = p->positive_weight; d.w_neg = p->negative_weight;
  d.n_obj = n_obj; d.n_lab = n_lab; d.n_class = ch - 7; d.hs = hs; d.ws = ws; d.ch = ch; d.nchw = nchw ? 1 : 0;
  d.step = (float)step;
  for (int i = 0; i < 3; ++i) d.rmax[i] = r_max[i];
  lp_loss_kernel<<<batch, kLossThreads, 0, (cudaStream_t)stream>>>(lp_map, labels, d, out_losses, batch, dlp);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}
