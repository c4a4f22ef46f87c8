/*
 * oracle.c -- CPU restatement of the Mask-RCNN-CoreML custom-layer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mask-rcnn-coreml_b200/ (the product)
 * may import, link or execute this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker / reported CPU baseline.
 *
 * PARITY UNPINNED: the reference (Swift + Core ML + Accelerate + MPS) cannot be
 * built or run on Linux and ships no tests, golden vectors or fixtures
 * (SURVEY.md section 4 and 8(c)).  This file is therefore a line-by-line
 * restatement of the Swift sources, in the "intended" mode of SURVEY.md
 * Appendix A (deterministic tie-breaks, reference bugs Q4/Q5 fixed).  Every
 * function cites the reference file:line it follows (paths are relative to
 * /root/reference/Sources/Mask-RCNN-CoreML/).
 *
 * Arithmetic contract (shared with the CUDA kernels):
 *   - fp32 box math, every operation individually rounded (no FMA contraction;
 *     compile with -ffp-contract=off), exp(Float) = (float)exp((double)x);
 *   - IoU in fp64 from exact fp32->fp64 conversions, quotient rounded to fp32,
 *     compared with a strict '>' against the fp32 threshold;
 *   - ROIAlign = TensorFlow crop_and_resize (bilinear, extrapolation 0), fp32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Utils.swift:17-26  stridedSlice: out[i*len+l] = p[begin + i*stride + l]    */
/* ------------------------------------------------------------------------- */
ORC_API void orc_strided_slice(const float* p, int begin, int count, int stride,
                               int length, float* out) {
  for (int l = 0; l < length; ++l)
    for (int i = 0; i < count; ++i)
      out[(size_t)i * length + l] = p[begin + (size_t)i * stride + l];
}

/* ------------------------------------------------------------------------- */
/* Utils.swift:56-66  sortedIndices(ascending:false) -- vDSP_vsorti.          */
/* Tie order is undocumented (Q1); intended mode = stable, lower index first. */
/* Bottom-up merge sort so that the full-N argsort the reference performs is  */
/* reproduced (and timed) faithfully.                                          */
/* ------------------------------------------------------------------------- */
static int desc_before(const float* key, uint32_t a, uint32_t b) {
  /* true when a must come before b: larger key first, ties by lower index */
  float ka = key[a], kb = key[b];
  if (ka > kb) return 1;
  if (ka < kb) return 0;
  return a < b;
}

ORC_API void orc_argsort_desc(const float* key, int n, uint32_t* idx) {
  uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) idx[i] = (uint32_t)i;
  uint32_t* src = idx;
  uint32_t* dst = tmp;
  for (int width = 1; width < n; width *= 2) {
    for (int lo = 0; lo < n; lo += 2 * width) {
      int mid = lo + width < n ? lo + width : n;
      int hi = lo + 2 * width < n ? lo + 2 * width : n;
      int i = lo, j = mid, k = lo;
      while (i < mid && j < hi) {
        if (desc_before(key, src[j], src[i])) dst[k++] = src[j++];
        else dst[k++] = src[i++];
      }
      while (i < mid) dst[k++] = src[i++];
      while (j < hi) dst[k++] = src[j++];
    }
    uint32_t* t = src; src = dst; dst = t;
  }
  if (src != idx) memcpy(idx, src, sizeof(uint32_t) * (size_t)n);
  free(tmp);
}

/* ------------------------------------------------------------------------- */
/* Utils.swift:173-180 elementWiseMultiply: M[r][c] *= v[c]  (rounded fp32)   */
/* ------------------------------------------------------------------------- */
ORC_API void orc_scale_rows(float* m, const float* v, int height, int width) {
  for (int c = 0; c < width; ++c)
    for (int r = 0; r < height; ++r) m[(size_t)r * width + c] = m[(size_t)r * width + c] * v[c];
}

/* ------------------------------------------------------------------------- */
/* BoxUtils.swift:32-71 applyBoxDeltas (in place; each op rounded to fp32)    */
/* ------------------------------------------------------------------------- */
static float expf_contract(float x) { return (float)exp((double)x); }

ORC_API void orc_apply_box_deltas(float* boxes, const float* deltas, int n) {
  for (int i = 0; i < n; ++i) {
    float* b = boxes + (size_t)i * 4;
    const float* d = deltas + (size_t)i * 4;
    float y1 = b[0], x1 = b[1], y2 = b[2], x2 = b[3];
    float height = y2 - y1;                 /* :50 */
    float width = x2 - x1;                  /* :51 */
    float hh = 0.5f * height;
    float hw = 0.5f * width;
    float centerY = y1 + hh;                /* :52 */
    float centerX = x1 + hw;                /* :53 */
    float ty = d[0] * height;
    float tx = d[1] * width;
    centerY = centerY + ty;                          /* :55 */
    centerX = centerX + tx;                          /* :56 */
    float eh = expf_contract(d[2]);
    float ew = expf_contract(d[3]);
    height = height * eh;                            /* :58 */
    width = width * ew;                              /* :59 */
    hh = 0.5f * height;
    hw = 0.5f * width;
    float ry1 = centerY - hh;               /* :61 */
    float rx1 = centerX - hw;               /* :62 */
    float ry2 = ry1 + height;               /* :63 */
    float rx2 = rx1 + width;                /* :64 */
    b[0] = ry1; b[1] = rx1; b[2] = ry2; b[3] = rx2;
  }
}

/* BoxUtils.swift:73-80 clip: vDSP_vclip to [0,1] */
ORC_API void orc_clip(float* v, int count) {
  for (int i = 0; i < count; ++i) {
    float x = v[i];
    if (x < 0.0f) x = 0.0f;
    if (x > 1.0f) x = 1.0f;
    v[i] = x;
  }
}

/* ------------------------------------------------------------------------- */
/* Utils.swift:222-229 CGRect(anchorDatum:) and Utils.swift:232-246 IOU.      */
/* CGFloat == Double on 64-bit Darwin. Boxes are (y1,x1,y2,x2) fp32.          */
/* CGRect.width / height / minX / maxX are those of the STANDARDISED rectangle  */
/* (|w|, |h|, min / max edge); for the boxes the layers decode themselves       */
/* (x2 >= x1, y2 >= y1) that is the identity.                                   */
/* ------------------------------------------------------------------------- */
ORC_API float orc_iou(const float* a, const float* b) {
  double ay1 = a[0], ax1 = a[1], ay2 = a[2], ax2 = a[3];
  double by1 = b[0], bx1 = b[1], by2 = b[2], bx2 = b[3];
  double aw = ax2 - ax1, ah = ay2 - ay1;
  double bw = bx2 - bx1, bh = by2 - by1;
  double areaA = fabs(aw * ah);
  if (areaA <= 0) return 0.0f;
  double areaB = fabs(bw * bh);
  if (areaB <= 0) return 0.0f;
  double aEndX = ax1 + aw, aEndY = ay1 + ah;
  double bEndX = bx1 + bw, bEndY = by1 + bh;
  double aMinX = aw >= 0 ? ax1 : aEndX, aMaxX = aw >= 0 ? aEndX : ax1;
  double aMinY = ah >= 0 ? ay1 : aEndY, aMaxY = ah >= 0 ? aEndY : ay1;
  double bMinX = bw >= 0 ? bx1 : bEndX, bMaxX = bw >= 0 ? bEndX : bx1;
  double bMinY = bh >= 0 ? by1 : bEndY, bMaxY = bh >= 0 ? bEndY : by1;
  double iMinX = aMinX > bMinX ? aMinX : bMinX;
  double iMinY = aMinY > bMinY ? aMinY : bMinY;
  double iMaxX = aMaxX < bMaxX ? aMaxX : bMaxX;
  double iMaxY = aMaxY < bMaxY ? aMaxY : bMaxY;
  double ih = iMaxY - iMinY;
  double iw = iMaxX - iMinX;
  if (ih < 0) ih = 0;
  if (iw < 0) iw = 0;
  double inter = ih * iw;
  double sum = areaA + areaB;
  double uni = sum - inter;
  double q = inter / uni;
  return (float)q;
}

/* ------------------------------------------------------------------------- */
/* Utils.swift:185-218 nonMaxSupression (greedy, given index order)           */
/* ------------------------------------------------------------------------- */
ORC_API int orc_nms(const float* boxes, const int* indices, int n_idx,
                    float iou_thr, int max_keep, int* selected) {
  int count = 0;
  for (int t = 0; t < n_idx; ++t) {
    if (count >= max_keep) return count;             /* :192 */
    int index = indices[t];
    const float* A = boxes + (size_t)index * 4;
    double w = (double)A[3] - (double)A[1];
    double h = (double)A[2] - (double)A[0];
    int keep = (fabs(w) > 0) && (fabs(h) > 0);        /* :195, CGRect.width / height = |w|, |h| */
    if (keep) {
      for (int s = 0; s < count; ++s) {              /* :200 */
        const float* B = boxes + (size_t)selected[s] * 4;
        if (orc_iou(A, B) > iou_thr) { keep = 0; break; } /* :203 */
      }
    }
    if (keep) selected[count++] = index;             /* :212 */
  }
  return count;
}

/* ------------------------------------------------------------------------- */
/* ProposalLayer.swift:103-195 evaluate                                        */
/*  probs (N,2), deltas (N,4), anchors (N,4) -> rois (max_prop,4) zero padded */
/*  keep_anchor (optional): anchor index of each kept roi, -1 padded.          */
/*  sorted_boxes_out (optional): the pre_n decoded+clipped boxes in score      */
/*  order (what nonMaxSupression sees).                                        */
/*  Returns number of proposals kept.                                          */
/* ------------------------------------------------------------------------- */
ORC_API int orc_proposal(const float* probs, const float* deltas,
                         const float* anchors, int N, const float* bbox_std,
                         int pre_nms_limit, int max_proposals, float iou_thr,
                         float* rois_out, int* keep_anchor,
                         float* sorted_boxes_out) {
  int n = N < pre_nms_limit ? N : pre_nms_limit;     /* :120 */
  float* score = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
  orc_strided_slice(probs, 1, N, 2, 1, score);       /* :124 */
  uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(N > 0 ? N : 1));
  orc_argsort_desc(score, N, order);                 /* :133 (full sort, then cut) */

  float* sd = (float*)malloc(sizeof(float) * 4 * (size_t)(n > 0 ? n : 1));
  float* sa = (float*)malloc(sizeof(float) * 4 * (size_t)(n > 0 ? n : 1));
  for (int k = 0; k < n; ++k) {                      /* :140-149 gathers */
    size_t a = order[k];
    for (int c = 0; c < 4; ++c) {
      sd[(size_t)k * 4 + c] = deltas[a * 4 + c];
      sa[(size_t)k * 4 + c] = anchors[a * 4 + c];
    }
  }
  orc_scale_rows(sd, bbox_std, n, 4);                /* :158 */
  orc_apply_box_deltas(sa, sd, n);                   /* :162 */
  orc_clip(sa, n * 4);                               /* :163 */

  /* :169-172; Q4: the reference passes 0..<4n as indices which traps when fewer
     than max survive; intended = 0..<n. */
  int* idx = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int k = 0; k < n; ++k) idx[k] = k;
  int* sel = (int*)malloc(sizeof(int) * (size_t)(max_proposals > 0 ? max_proposals : 1));
  int cnt = orc_nms(sa, idx, n, iou_thr, max_proposals, sel);

  for (int i = 0; i < cnt; ++i) {                    /* :181-185 */
    for (int j = 0; j < 4; ++j) rois_out[(size_t)i * 4 + j] = sa[(size_t)sel[i] * 4 + j];
    if (keep_anchor) keep_anchor[i] = (int)order[sel[i]];
  }
  for (int i = cnt; i < max_proposals; ++i) {        /* :190-192 */
    for (int j = 0; j < 4; ++j) rois_out[(size_t)i * 4 + j] = 0.0f;
    if (keep_anchor) keep_anchor[i] = -1;
  }
  if (sorted_boxes_out) memcpy(sorted_boxes_out, sa, sizeof(float) * 4 * (size_t)n);
  free(score); free(order); free(sd); free(sa); free(idx); free(sel);
  return cnt;
}

/* ------------------------------------------------------------------------- */
/* PyramidROIAlignLayer.swift:351-396 roisToInputItems: level assignment.      */
/* level_out[i] in {2..5}, or -1 for a padding (invalid) roi.                  */
/* ------------------------------------------------------------------------- */
ORC_API void orc_roi_levels(const float* rois, int roi_stride, int R,
                            double image_w, double image_h, int* level_out) {
  double ratio = 224.0 / sqrt(image_w * image_h);    /* :357 (factor :98) */
  for (int i = 0; i < R; ++i) {
    const float* r = rois + (size_t)i * roi_stride;
    double y1 = r[0], x1 = r[1], y2 = r[2], x2 = r[3];
    double width = x2 - x1, height = y2 - y1;
    double lf = log2(sqrt(width * height) / ratio) + 4.0;   /* :373 */
    int valid = !isnan(lf) && !isinf(lf);            /* :374 */
    if (!valid) { level_out[i] = -1; continue; }
    double rl = round(lf);                           /* half away from zero, Q13 */
    int lvl = (int)rl;
    if (lvl < 2) lvl = 2;
    if (lvl > 5) lvl = 5;                            /* :376 */
    level_out[i] = lvl;
  }
}

/* ------------------------------------------------------------------------- */
/* MPSNNCropAndResizeBilinear (PyramidROIAlignLayer.swift:212-223) restated as */
/* TensorFlow crop_and_resize, bilinear, extrapolation value 0, fp32.          */
/* One region on one CHW map -> (C,P,P).                                       */
/* ------------------------------------------------------------------------- */
static void crop_and_resize_chw(const float* fmap, int C, int H, int W,
                                float y1, float x1, float y2, float x2, int P,
                                float* out) {
  float hm1 = (float)(H - 1), wm1 = (float)(W - 1);
  float hs = 0.0f, ws = 0.0f;
  if (P > 1) {
    float dy = y2 - y1, dx = x2 - x1;
    float ny = dy * hm1, nx = dx * wm1;
    hs = ny / (float)(P - 1);
    ws = nx / (float)(P - 1);
  }
  for (int py = 0; py < P; ++py) {
    float in_y;
    if (P > 1) { float a = y1 * hm1; float b = (float)py * hs; in_y = a + b; }
    else { float s = y1 + y2; float m = 0.5f * s; in_y = m * hm1; }
    int y_ok = !(in_y < 0.0f || in_y > hm1);
    float fy = floorf(in_y), cy = ceilf(in_y);
    int t = (int)fy, b = (int)cy;
    float ly = in_y - fy;
    for (int px = 0; px < P; ++px) {
      float in_x;
      if (P > 1) { float a = x1 * wm1; float bb = (float)px * ws; in_x = a + bb; }
      else { float s = x1 + x2; float m = 0.5f * s; in_x = m * wm1; }
      int x_ok = !(in_x < 0.0f || in_x > wm1);
      if (!y_ok || !x_ok) {
        for (int c = 0; c < C; ++c) out[((size_t)c * P + py) * P + px] = 0.0f;
        continue;
      }
      float fx = floorf(in_x), cx = ceilf(in_x);
      int l = (int)fx, r = (int)cx;
      float lx = in_x - fx;
      for (int c = 0; c < C; ++c) {
        const float* pl = fmap + (size_t)c * H * W;
        float tl = pl[(size_t)t * W + l], tr = pl[(size_t)t * W + r];
        float bl = pl[(size_t)b * W + l], br = pl[(size_t)b * W + r];
        float dt = tr - tl; float mt = dt * lx; float top = tl + mt;
        float db = br - bl; float mb = db * lx; float bot = bl + mb;
        float dv = bot - top; float mv = dv * ly; float o = top + mv;
        out[((size_t)c * P + py) * P + px] = o;
      }
    }
  }
}

/* ------------------------------------------------------------------------- */
/* PyramidROIAlignLayer.swift:79-181 evaluate (+ :245-274 copyOutput).        */
/* fmaps: 4 CHW maps (levels 2..5), hw[l] = {H_l, W_l}.  out (R,C,P,P).        */
/* Q5 fixed: every block is written; invalid rois produce zero blocks.         */
/* ------------------------------------------------------------------------- */
ORC_API void orc_pyramid_roialign(const float* rois, int roi_stride, int R,
                                  const float* f2, const float* f3,
                                  const float* f4, const float* f5,
                                  const int* hw, int C, int P, double image_w,
                                  double image_h, float* out, int* level_out) {
  const float* maps[4] = {f2, f3, f4, f5};
  int* lv = (int*)malloc(sizeof(int) * (size_t)(R > 0 ? R : 1));
  orc_roi_levels(rois, roi_stride, R, image_w, image_h, lv);
  size_t blk = (size_t)C * P * P;
  for (int i = 0; i < R; ++i) {
    float* o = out + (size_t)i * blk;
    if (lv[i] < 0) { memset(o, 0, sizeof(float) * blk); continue; }
    int m = lv[i] - 2;
    const float* r = rois + (size_t)i * roi_stride;
    crop_and_resize_chw(maps[m], C, hw[2 * m], hw[2 * m + 1], r[0], r[1], r[2], r[3], P, o);
  }
  if (level_out) memcpy(level_out, lv, sizeof(int) * (size_t)R);
  free(lv);
}

/* ------------------------------------------------------------------------- */
/* TimeDistributedClassifierLayer.swift:50-88 + :177-192: first-max argmax,    */
/* pick that class's 4 deltas (class-major (ncls,4), :83).                     */
/* probs (R,ncls), bbox (R,ncls*4) -> out (R,6) = [d0..d3, classId, score]     */
/* ------------------------------------------------------------------------- */
ORC_API void orc_classifier_select(const float* probs, const float* bbox, int R,
                                   int ncls, float* out) {
  for (int r = 0; r < R; ++r) {
    const float* p = probs + (size_t)r * ncls;
    int best = 0; float bv = p[0];
    for (int c = 1; c < ncls; ++c) if (p[c] > bv) { bv = p[c]; best = c; }  /* vDSP_maxvi: first max */
    const float* d = bbox + ((size_t)r * ncls + best) * 4;
    float* o = out + (size_t)r * 6;
    o[0] = d[0]; o[1] = d[1]; o[2] = d[2]; o[3] = d[3];
    o[4] = (float)best; o[5] = bv;
  }
}

/* ------------------------------------------------------------------------- */
/* DetectionLayer.swift:107-234 evaluate (+ :238-276).                          */
/* rois (R,4), cls (R,6) -> out (max_det,6) zero padded; keep_roi optional      */
/* (roi index of each output row, -1 padded). Returns detection count.          */
/* Intended-mode orders: classes ascending (Q11), final sort stable (Q12).      */
/* ------------------------------------------------------------------------- */
ORC_API int orc_detection(const float* rois, const float* cls, int R,
                          const float* bbox_std, int max_det, float score_thr,
                          float iou_thr, float* out, int* keep_roi) {
  int* fidx = (int*)malloc(sizeof(int) * (size_t)(R > 0 ? R : 1));
  int K = 0;
  for (int r = 0; r < R; ++r) {
    float score = cls[(size_t)r * 6 + 5];
    float classId = cls[(size_t)r * 6 + 4];
    /* :259 vthres keeps score >= thr, :267 vcmprs keeps non-zero gate; :136-140 classId > 0 */
    if (score >= score_thr && score != 0.0f && classId > 0.0f) fidx[K++] = r;
  }
  float* fr = (float*)malloc(sizeof(float) * 4 * (size_t)(K > 0 ? K : 1));
  float* fd = (float*)malloc(sizeof(float) * 4 * (size_t)(K > 0 ? K : 1));
  float* fs = (float*)malloc(sizeof(float) * (size_t)(K > 0 ? K : 1));
  float* fc = (float*)malloc(sizeof(float) * (size_t)(K > 0 ? K : 1));
  for (int k = 0; k < K; ++k) {                      /* :144-154 */
    int r = fidx[k];
    for (int c = 0; c < 4; ++c) {
      fr[(size_t)k * 4 + c] = rois[(size_t)r * 4 + c];
      fd[(size_t)k * 4 + c] = cls[(size_t)r * 6 + c];
    }
    fs[k] = cls[(size_t)r * 6 + 5];
    fc[k] = cls[(size_t)r * 6 + 4];
  }
  orc_scale_rows(fd, bbox_std, K, 4);                /* :159 */
  orc_apply_box_deltas(fr, fd, K);                   /* :163 */
  orc_clip(fr, K * 4);                               /* :164 */

  /* :166-183 per-class NMS; classes visited ascending (intended) */
  int* nms_ids = (int*)malloc(sizeof(int) * (size_t)(K > 0 ? K : 1));
  int n_ids = 0;
  int* of_class = (int*)malloc(sizeof(int) * (size_t)(K > 0 ? K : 1));
  int* sel = (int*)malloc(sizeof(int) * (size_t)(max_det > 0 ? max_det : 1));
  float last = -INFINITY;
  for (;;) {
    /* next distinct class value greater than 'last' */
    float next = INFINITY; int found = 0;
    for (int k = 0; k < K; ++k) if (fc[k] > last && fc[k] < next) { next = fc[k]; found = 1; }
    if (!found) break;
    last = next;
    int m = 0;
    for (int k = 0; k < K; ++k) if (fc[k] == next) of_class[m++] = k;   /* :172-176 */
    int c = orc_nms(fr, of_class, m, iou_thr, max_det, sel);            /* :178-181 */
    for (int i = 0; i < c; ++i) nms_ids[n_ids++] = sel[i];
  }
  /* :186-209 sort by score desc (stable), keep first min(count,max_det) */
  for (int i = 1; i < n_ids; ++i) {                  /* insertion sort = stable */
    int v = nms_ids[i]; int j = i - 1;
    while (j >= 0 && fs[nms_ids[j]] < fs[v]) { nms_ids[j + 1] = nms_ids[j]; --j; }
    nms_ids[j + 1] = v;
  }
  int cnt = n_ids < max_det ? n_ids : max_det;
  for (int i = 0; i < cnt; ++i) {                    /* :217-224 */
    int k = nms_ids[i];
    float* o = out + (size_t)i * 6;
    o[0] = fr[(size_t)k * 4]; o[1] = fr[(size_t)k * 4 + 1];
    o[2] = fr[(size_t)k * 4 + 2]; o[3] = fr[(size_t)k * 4 + 3];
    o[4] = fc[k]; o[5] = fs[k];
    if (keep_roi) keep_roi[i] = fidx[k];
  }
  for (int i = cnt; i < max_det; ++i) {              /* :228-231 */
    for (int j = 0; j < 6; ++j) out[(size_t)i * 6 + j] = 0.0f;
    if (keep_roi) keep_roi[i] = -1;
  }
  free(fidx); free(fr); free(fd); free(fs); free(fc); free(nms_ids); free(of_class); free(sel);
  return cnt;
}

/* ------------------------------------------------------------------------- */
/* TimeDistributedMaskLayer.swift:58-89: class-plane selection.                 */
/* masks_all (D,ncls,S,S) is what Mask.mlmodel returns for every slot;          */
/* valid[i] says whether pooled block i had no zero element (Q9, intended =     */
/* "roi i is not padding").  out (D,S,S).                                       */
/* ------------------------------------------------------------------------- */
ORC_API void orc_mask_select(const float* masks_all, const int* valid,
                             const float* detections, int D, int ncls, int S,
                             float* out) {
  size_t plane = (size_t)S * S;
  int j = 0;
  memset(out, 0, sizeof(float) * plane * (size_t)D);
  for (int a = 0; a < D; ++a) {
    if (!valid[a]) continue;
    int cls = (int)detections[(size_t)j * 6 + 4];    /* :71 indexed by compacted j (Q14) */
    if (cls < 0) cls = 0;
    if (cls >= ncls) cls = ncls - 1;
    memcpy(out + (size_t)a * plane, masks_all + ((size_t)a * ncls + cls) * plane,
           sizeof(float) * plane);                   /* :83 */
    ++j;
  }
  /* :87-89 rows [count(valid), D) are zeroed */
  for (int a = j; a < D; ++a) memset(out + (size_t)a * plane, 0, sizeof(float) * plane);
}

/* ------------------------------------------------------------------------- */
/* Detection.swift:23-62 + :64-99: public decoding of (D,6) + (D,S,S).          */
/* keep row iff Double(score) > 0.7; bbox = (x1,y1,w,h) in Double;               */
/* mask byte = UInt8(255 - p/2*255) (truncation).                                */
/* Returns count; index_out[i], bbox_out[i*4..] (x,y,w,h), class_out, score_out, */
/* mask_out (count,S*S) bytes.                                                   */
/* ------------------------------------------------------------------------- */
ORC_API int orc_detections_decode(const float* det, const float* masks, int D,
                                  int S, int* index_out, double* bbox_out,
                                  int* class_out, double* score_out,
                                  uint8_t* mask_out) {
  int n = 0;
  size_t plane = (size_t)S * S;
  for (int i = 0; i < D; ++i) {
    double score = (double)det[(size_t)i * 6 + 5];
    if (!(score > 0.7)) continue;                    /* :38 */
    double y1 = det[(size_t)i * 6], x1 = det[(size_t)i * 6 + 1];
    double y2 = det[(size_t)i * 6 + 2], x2 = det[(size_t)i * 6 + 3];
    index_out[n] = i;
    bbox_out[(size_t)n * 4] = x1; bbox_out[(size_t)n * 4 + 1] = y1;
    bbox_out[(size_t)n * 4 + 2] = x2 - x1; bbox_out[(size_t)n * 4 + 3] = y2 - y1;
    class_out[n] = (int)det[(size_t)i * 6 + 4];
    score_out[n] = score;
    if (masks && mask_out) {
      for (size_t p = 0; p < plane; ++p) {
        double v = (double)masks[(size_t)i * plane + p];
        double b = 255.0 - (v / 2.0 * 255.0);        /* :84 */
        mask_out[(size_t)n * plane + p] = (uint8_t)b;
      }
    }
    ++n;
  }
  return n;
}

/* ------------------------------------------------------------------------- */
/* Internal-layout variant of the same crop_and_resize used inside the fused   */
/* pipeline: feature maps are NHWC fp16 (H,W,C), output is (R,P,P,C) fp16.      */
/* Same fp32 arithmetic as above on the fp16->fp32 converted taps, result       */
/* rounded to fp16 (round-to-nearest-even).  Levels as orc_roi_levels.          */
/* ------------------------------------------------------------------------- */
typedef _Float16 orc_half;

static void crop_and_resize_hwc_f16(const orc_half* fmap, int C, int H, int W,
                                    float y1, float x1, float y2, float x2,
                                    int P, orc_half* out) {
  float hm1 = (float)(H - 1), wm1 = (float)(W - 1);
  float hs = 0.0f, ws = 0.0f;
  if (P > 1) {
    float dy = y2 - y1, dx = x2 - x1;
    float ny = dy * hm1, nx = dx * wm1;
    hs = ny / (float)(P - 1);
    ws = nx / (float)(P - 1);
  }
  for (int py = 0; py < P; ++py) {
    float in_y;
    if (P > 1) { float a = y1 * hm1; float b = (float)py * hs; in_y = a + b; }
    else { float s = y1 + y2; float m = 0.5f * s; in_y = m * hm1; }
    int y_ok = !(in_y < 0.0f || in_y > hm1);
    float fy = floorf(in_y), cy = ceilf(in_y);
    int t = (int)fy, b = (int)cy;
    float ly = in_y - fy;
    for (int px = 0; px < P; ++px) {
      float in_x;
      if (P > 1) { float a = x1 * wm1; float bb = (float)px * ws; in_x = a + bb; }
      else { float s = x1 + x2; float m = 0.5f * s; in_x = m * wm1; }
      int x_ok = !(in_x < 0.0f || in_x > wm1);
      orc_half* o = out + ((size_t)py * P + px) * C;
      if (!y_ok || !x_ok) { for (int c = 0; c < C; ++c) o[c] = (orc_half)0.0f; continue; }
      float fx = floorf(in_x), cx = ceilf(in_x);
      int l = (int)fx, r = (int)cx;
      float lx = in_x - fx;
      const orc_half* ptl = fmap + ((size_t)t * W + l) * C;
      const orc_half* ptr_ = fmap + ((size_t)t * W + r) * C;
      const orc_half* pbl = fmap + ((size_t)b * W + l) * C;
      const orc_half* pbr = fmap + ((size_t)b * W + r) * C;
      for (int c = 0; c < C; ++c) {
        float tl = (float)ptl[c], tr = (float)ptr_[c], bl = (float)pbl[c], br = (float)pbr[c];
        float dt = tr - tl; float mt = dt * lx; float top = tl + mt;
        float db = br - bl; float mb = db * lx; float bot = bl + mb;
        float dv = bot - top; float mv = dv * ly; float ov = top + mv;
        o[c] = (orc_half)ov;
      }
    }
  }
}

ORC_API void orc_pyramid_roialign_nhwc_f16(const float* rois, int roi_stride, int R,
                                           const uint16_t* f2, const uint16_t* f3,
                                           const uint16_t* f4, const uint16_t* f5,
                                           const int* hw, int C, int P,
                                           double image_w, double image_h,
                                           uint16_t* out, int* level_out) {
  const orc_half* maps[4] = {(const orc_half*)f2, (const orc_half*)f3,
                             (const orc_half*)f4, (const orc_half*)f5};
  int* lv = (int*)malloc(sizeof(int) * (size_t)(R > 0 ? R : 1));
  orc_roi_levels(rois, roi_stride, R, image_w, image_h, lv);
  size_t blk = (size_t)C * P * P;
  for (int i = 0; i < R; ++i) {
    orc_half* o = (orc_half*)out + (size_t)i * blk;
    if (lv[i] < 0) { memset(o, 0, sizeof(orc_half) * blk); continue; }
    int m = lv[i] - 2;
    const float* r = rois + (size_t)i * roi_stride;
    crop_and_resize_hwc_f16(maps[m], C, hw[2 * m], hw[2 * m + 1], r[0], r[1], r[2], r[3], P, o);
  }
  if (level_out) memcpy(level_out, lv, sizeof(int) * (size_t)R);
  free(lv);
}

/* ------------------------------------------------------------------------- */
/* Letter-boxing in front of the path (Vision .scaleFit, EvaluateCommand.swift:157).  Geometry =                */
/* DetectionRenderer.swift:63-75 (scale factor by fitsHorizontally, padding split evenly); bilinear with        */
/* half-pixel centres (Vision's filter is closed: PARITY UNPINNED), fp64, black padding, round half up to u8.   */
/* ------------------------------------------------------------------------- */
ORC_API void orc_letterbox(const uint8_t* src, int src_h, int src_w, int dst_h, int dst_w, uint8_t* dst) {
  double hs = (double)dst_w / (double)src_w, vs = (double)dst_h / (double)src_h;      /* :63-64 */
  int fits_h = (double)src_h * hs <= (double)dst_h;                                   /* :66 */
  double scale = fits_h ? hs : vs;                                                    /* :68 */
  double new_w = (double)src_w * scale, new_h = (double)src_h * scale;                /* :70 */
  double pad_x = ((double)dst_w - new_w) / 2.0, pad_y = ((double)dst_h - new_h) / 2.0;/* :72-75 */
  for (int Y = 0; Y < dst_h; ++Y)
    for (int X = 0; X < dst_w; ++X) {
      uint8_t* o = dst + ((size_t)Y * dst_w + X) * 3;
      double cx = ((double)X + 0.5) - pad_x, cy = ((double)Y + 0.5) - pad_y;
      if (!(cx >= 0.0 && cx < new_w && cy >= 0.0 && cy < new_h)) { o[0] = o[1] = o[2] = 0; continue; }
      double sx = cx / scale - 0.5, sy = cy / scale - 0.5;
      if (sx < 0.0) sx = 0.0;
      if (sx > (double)(src_w - 1)) sx = (double)(src_w - 1);
      if (sy < 0.0) sy = 0.0;
      if (sy > (double)(src_h - 1)) sy = (double)(src_h - 1);
      int x0 = (int)floor(sx), y0 = (int)floor(sy);
      int x1 = x0 + 1 < src_w ? x0 + 1 : x0, y1 = y0 + 1 < src_h ? y0 + 1 : y0;
      double fx = sx - (double)x0, fy = sy - (double)y0;
      for (int c = 0; c < 3; ++c) {
        double tl = src[((size_t)y0 * src_w + x0) * 3 + c], tr = src[((size_t)y0 * src_w + x1) * 3 + c];
        double bl = src[((size_t)y1 * src_w + x0) * 3 + c], br = src[((size_t)y1 * src_w + x1) * 3 + c];
        double dt = tr - tl; double mt = dt * fx; double top = tl + mt;
        double db = br - bl; double mb = db * fx; double bot = bl + mb;
        double dv = bot - top; double mv = dv * fy; double v = top + mv;
        double r = v + 0.5;
        o[c] = (uint8_t)r;
      }
    }
}
