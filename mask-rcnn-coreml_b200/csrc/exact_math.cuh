// exact_math.cuh -- the arithmetic contract shared with oracle/oracle.c.
// Every operation is individually rounded (explicit _rn intrinsics, so the
// result does not depend on -fmad); see DESIGN.md "Arithmetic contract".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Order-preserving float -> uint key (larger score -> larger key).
// -0 is canonicalised to +0 so that it ties with +0 like a float compare does;
// NaN maps to 0 (sorted last); the oracle's comparator is undefined on NaN.
__device__ __forceinline__ uint32_t score_key(float s) {
  if (s != s) return 0u;
  s = __fadd_rn(s, 0.0f);
  uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float clip01(float x) {
  // BoxUtils.swift:73-80 (vDSP_vclip); comparisons keep NaN like the oracle.
  x = (x < 0.0f) ? 0.0f : x;
  x = (x > 1.0f) ? 1.0f : x;
  return x;
}

// ProposalLayer.swift:158 / DetectionLayer.swift:159 (x std, rounded) followed by
// BoxUtils.swift:32-71 (applyBoxDeltas) and :73-80 (clip).
// box = (y1,x1,y2,x2), d = (dy,dx,log dh,log dw).
__device__ __forceinline__ float4 decode_box(float4 box, float4 d, float4 sd) {
  float d0 = __fmul_rn(d.x, sd.x), d1 = __fmul_rn(d.y, sd.y);
  float d2 = __fmul_rn(d.z, sd.z), d3 = __fmul_rn(d.w, sd.w);
  float y1 = box.x, x1 = box.y, y2 = box.z, x2 = box.w;
  float height = __fsub_rn(y2, y1);
  float width = __fsub_rn(x2, x1);
  float cy = __fadd_rn(y1, __fmul_rn(0.5f, height));
  float cx = __fadd_rn(x1, __fmul_rn(0.5f, width));
  cy = __fadd_rn(cy, __fmul_rn(d0, height));
  cx = __fadd_rn(cx, __fmul_rn(d1, width));
  float eh = (float)exp((double)d2);
  float ew = (float)exp((double)d3);
  height = __fmul_rn(height, eh);
  width = __fmul_rn(width, ew);
  float ry1 = __fsub_rn(cy, __fmul_rn(0.5f, height));
  float rx1 = __fsub_rn(cx, __fmul_rn(0.5f, width));
  float ry2 = __fadd_rn(ry1, height);
  float rx2 = __fadd_rn(rx1, width);
  return make_float4(clip01(ry1), clip01(rx1), clip01(ry2), clip01(rx2));
}

// Utils.swift:222-229: CGRect(x: x1, y: y1, width: x2-x1, height: y2-y1) in Double.
struct RectD {
  double x1, y1, maxx, maxy, area;
};

__device__ __forceinline__ RectD make_rect(float4 b) {
  // CGRect(x: x1, y: y1, width: x2 - x1, height: y2 - y1) (Utils.swift:222-231); CGRect's width / height / minX / maxX are
  // those of the STANDARDISED rectangle: |w|, |h|, min and max edge.  Boxes with x2 >= x1, y2 >= y1 (every box the
  // layers decode themselves) give exactly the values of the plain formulas.
  RectD r;
  double y1 = (double)b.x, x1 = (double)b.y, y2 = (double)b.z, x2 = (double)b.w;
  double w = __dsub_rn(x2, x1), h = __dsub_rn(y2, y1);
  const double xe = __dadd_rn(x1, w), ye = __dadd_rn(y1, h);
  r.x1 = w >= 0.0 ? x1 : xe; r.maxx = w >= 0.0 ? xe : x1;
  r.y1 = h >= 0.0 ? y1 : ye; r.maxy = h >= 0.0 ? ye : y1;
  r.area = fabs(__dmul_rn(w, h));
  return r;
}

__device__ __forceinline__ bool box_selectable(float4 b) {
  // Utils.swift:195: anchorA.width > 0 && anchorA.height > 0 (Double; width / height of the standardised rect = |w|, |h|)
  double w = __dsub_rn((double)b.w, (double)b.y), h = __dsub_rn((double)b.z, (double)b.x);
  return (fabs(w) > 0.0) && (fabs(h) > 0.0);
}

// x2 >= x1 and y2 >= y1: the box is its own standardised rectangle (the fp32 fast path of nms_mask_kernel needs that)
__device__ __forceinline__ bool box_ordered(float4 b) { return b.w >= b.y && b.z >= b.x; }

// Utils.swift:232-246 IOU: Double math, quotient rounded to Float.
__device__ __forceinline__ float iou_rect(const RectD& a, const RectD& b) {
  if (a.area <= 0.0) return 0.0f;
  if (b.area <= 0.0) return 0.0f;
  double ix0 = a.x1 > b.x1 ? a.x1 : b.x1;
  double iy0 = a.y1 > b.y1 ? a.y1 : b.y1;
  double ix1 = a.maxx < b.maxx ? a.maxx : b.maxx;
  double iy1 = a.maxy < b.maxy ? a.maxy : b.maxy;
  double ih = __dsub_rn(iy1, iy0), iw = __dsub_rn(ix1, ix0);
  // disjoint boxes: inter = 0 -> 0 / union = +0 exactly; skip the fp64 multiply / divide (most pairs)
  if (!(ih > 0.0) || !(iw > 0.0)) { if (!(ih != ih) && !(iw != iw)) return 0.0f; }
  ih = ih < 0.0 ? 0.0 : ih;
  iw = iw < 0.0 ? 0.0 : iw;
  double inter = __dmul_rn(ih, iw);
  double uni = __dsub_rn(__dadd_rn(a.area, b.area), inter);
  return __double2float_rn(__ddiv_rn(inter, uni));
}
