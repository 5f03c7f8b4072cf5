// Geometry of the detection heads shared by the decode and the training-target kernels.
#pragma once
#include "common.cuh"

namespace yb {

struct DecodeDev {
  const float* head[YOLO_MAX_SCALES];
  int boxes[YOLO_MAX_SCALES];      // cells * A per scale
  int ws[YOLO_MAX_SCALES];         // cells per row
  int off[YOLO_MAX_SCALES + 1];    // prefix sum of boxes
  float step[YOLO_MAX_SCALES];
  float anc[YOLO_MAX_SCALES][YOLO_MAX_ANCHORS][2];
  int n_scales, A, C, total;
  float img_h, img_w;
};


// Validates the C-ABI geometry and fills the device-side copy (heads may be null pointers when `need_heads` is false).
inline int make_dev(const yolo_decode_geom* g, const void* const* heads, DecodeDev& d) {
  if (!g || !heads) return fail(YOLO_E_BADARG, "decode: null geometry or heads");
  if (g->n_scales < 1 || g->n_scales > YOLO_MAX_SCALES || g->n_anchors < 1 || g->n_anchors > YOLO_MAX_ANCHORS)
    return fail(YOLO_E_BADARG, "decode: n_scales=%d n_anchors=%d out of range", g->n_scales, g->n_anchors);
  if (g->channels_per_anchor < 6) return fail(YOLO_E_BADARG, "decode: channels_per_anchor=%d < 6", g->channels_per_anchor);
  d.n_scales = g->n_scales; d.A = g->n_anchors; d.C = g->channels_per_anchor;
  d.img_h = (float)g->height; d.img_w = (float)g->width;
  d.off[0] = 0;
  for (int s = 0; s < g->n_scales; ++s) {
    if (g->step[s] <= 0 || g->height % g->step[s] || g->width % g->step[s])
      return fail(YOLO_E_SHAPE, "decode: size %dx%d not divisible by step %d", g->height, g->width, g->step[s]);
    if (!heads[s]) return fail(YOLO_E_BADARG, "decode: heads[%d] is null", s);
    d.head[s] = static_cast<const float*>(heads[s]);
    d.ws[s] = g->width / g->step[s];
    d.boxes[s] = (g->height / g->step[s]) * d.ws[s] * g->n_anchors;
    d.off[s + 1] = d.off[s] + d.boxes[s];
    d.step[s] = (float)g->step[s];
    for (int a = 0; a < g->n_anchors; ++a) { d.anc[s][a][0] = g->anchors[s][a][0]; d.anc[s][a][1] = g->anchors[s][a][1]; }
  }
  for (int s = g->n_scales; s < YOLO_MAX_SCALES; ++s) { d.head[s] = nullptr; d.boxes[s] = 0; d.ws[s] = 1; d.off[s + 1] = d.off[s]; d.step[s] = 1.f; }
  d.total = d.off[g->n_scales];
  return YOLO_OK;
}


}  // namespace yb
