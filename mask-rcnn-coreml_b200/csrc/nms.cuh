// nms.cuh -- greedy NMS (Utils.swift:185-218) as two device stages:
//   1. nms_mask_kernel: upper-triangular 64x64-tile suppression bitmask.  The reference's test is
//      Float(IoU in Double) > thr (Utils.swift:203,232-246); pairs whose fp32 IoU is further than 1e-4 from the
//      threshold are decided in fp32 (two orders of magnitude above its rounding error), borderline pairs with the
//      exact fp64 formula, so the bitmask equals the reference's decisions bit for bit;
//   2. nms_resolve(): chunked sequential scan that reproduces the reference's
//      visiting order exactly (box k is dropped iff an earlier KEPT box overlaps
//      it), with an optional total cap (ProposalLayer: maxProposals) and an
//      optional per-class cap (DetectionLayer: maxDetections per class).
// Greedy NMS is order dependent, so no "fast NMS" approximations are used.
#pragma once
#include "exact_math.cuh"

#define NMS_TILE 64

// boxes: [batch][stride_boxes] float4 (y1,x1,y2,x2); cls (optional): [batch][stride_boxes]
// counts (optional): per image number of boxes, else n_fixed.
// mask: [batch][stride_boxes][words] ; word (i, w) holds bits for columns 64w..64w+63,
// only words with w >= i/64 are written/read.
static __global__ void __launch_bounds__(NMS_TILE)
nms_mask_kernel(const float4* __restrict__ boxes, const float* __restrict__ cls,
                const int32_t* __restrict__ counts, int n_fixed, int stride_boxes,
                int words, float thr, unsigned long long* __restrict__ mask) {
  const int img = blockIdx.z;
  const int n = counts ? counts[img] : n_fixed;
  const int cb = blockIdx.x, rb = blockIdx.y;
  if (cb < rb) return;
  if (rb * NMS_TILE >= n || cb * NMS_TILE >= n) return;
  const float4* b = boxes + (size_t)img * stride_boxes;
  const float* c = cls ? cls + (size_t)img * stride_boxes : nullptr;

  __shared__ RectD col_rect[NMS_TILE];
  __shared__ float4 col_f[NMS_TILE];      // (lb(minx), lb(miny), ub(maxx), ub(maxy)) in fp32: lower bounds rounded down, upper up
  __shared__ float col_cls[NMS_TILE];
  __shared__ float4 col_box[NMS_TILE];    // (y1, x1, y2, x2) as given
  __shared__ float col_area[NMS_TILE];    // fp32 area estimate (sign-exact: zero / negative exactly when the fp64 area is)
  const int t = threadIdx.x;
  const int cj = cb * NMS_TILE + t;
  if (cj < n) {
    const RectD rc = make_rect(b[cj]);
    col_rect[t] = rc;
    col_box[t] = b[cj];
    // inverted boxes (caller-supplied anchors / rois only) never take the fp32 path: area estimate 0
    col_area[t] = box_ordered(b[cj]) ? (b[cj].w - b[cj].y) * (b[cj].z - b[cj].x) : 0.0f;
    col_f[t] = make_float4(__double2float_rd(rc.x1), __double2float_rd(rc.y1), __double2float_ru(rc.maxx), __double2float_ru(rc.maxy));
    col_cls[t] = c ? c[cj] : 0.0f;
  }
  __syncthreads();
  const int i = rb * NMS_TILE + t;
  if (i >= n) return;
  const RectD me = make_rect(b[i]);
  const float mx1 = __double2float_rd(me.x1), my1 = __double2float_rd(me.y1);      // lower bounds of the min edges
  const float mubx = __double2float_ru(me.maxx), muby = __double2float_ru(me.maxy);
  const float mycls = c ? c[i] : 0.0f;
  const float4 mb = b[i];
  const float marea = box_ordered(mb) ? (mb.w - mb.y) * (mb.z - mb.x) : 0.0f;
  const int ncol = min(NMS_TILE, n - cb * NMS_TILE);
  const bool prefilter = thr >= 0.0f;     // IoU of a disjoint pair is exactly 0, never > a non-negative threshold
  unsigned long long bits = 0ull;
  const int j0 = (cb == rb) ? t + 1 : 0;
  for (int j = j0; j < ncol; ++j) {
    if (c && col_cls[j] != mycls) continue;
    // fp32 pre-test with conservative bounds: ub >= max edge exactly, so "ub <= other's min edge" proves the boxes
    // do not overlap (intersection width or height <= 0 -> IoU 0 in the exact fp64 formula as well)
    const float4 cf = col_f[j];
    if (prefilter && ((cf.z <= mx1) || (mubx <= cf.x) || (cf.w <= my1) || (muby <= cf.y))) continue;
    // overlapping pair: decide in fp32 when the fp32 IoU is further than 1e-4 from the threshold (its error against
    // the exact formula is below 1e-6: a handful of roundings, and union >= max(area) so nothing cancels); only
    // borderline pairs take the exact path (fp64 IoU rounded to fp32, Utils.swift:232-246)
    const float4 cb4 = col_box[j];
    const float iw = fminf(mb.w, cb4.w) - fmaxf(mb.y, cb4.y), ih = fminf(mb.z, cb4.z) - fmaxf(mb.x, cb4.x);
    const float inter = fmaxf(iw, 0.0f) * fmaxf(ih, 0.0f);
    const float i32 = inter / (marea + col_area[j] - inter);
    bool sup;
    if (prefilter && marea > 0.0f && col_area[j] > 0.0f && fabsf(i32 - thr) > 1e-4f) sup = i32 > thr;
    else sup = iou_rect(col_rect[j], me) > thr;
    if (sup) bits |= (1ull << j);
  }
  mask[((size_t)img * stride_boxes + i) * words + cb] = bits;
}

// Sequential resolution, executed by one CTA (blockDim.x >= 256, multiple of 32).  Boxes are visited in
// super-chunks of NMS_SUPER mask words (256 boxes): everything the order-dependent walk needs (the intra-chunk
// block of the bitmask, selectability bits, classes) is first staged in shared memory by all threads, one thread
// then walks the alive candidates touching shared memory only, and all threads OR the kept rows into the removal
// words of the later chunks.  Returns (in every thread) the number of kept boxes; kept indices are appended to
// kept_out (shared or global) in visiting order.
#define NMS_SUPER 4
#define NMS_SUPER_BOXES (NMS_SUPER * NMS_TILE)

struct NmsResolveSmem {
  unsigned long long* remv;       // [words]
  unsigned long long* diag;       // [NMS_SUPER_BOXES * NMS_SUPER]
  unsigned long long* svalid;     // [NMS_SUPER]
  float* scls;                    // [NMS_SUPER_BOXES]
  int* kept_rows;                 // [NMS_SUPER_BOXES]
  int* misc;                      // [4]: 0 kept_total, 1 kept_in_chunk, 2 stop flag
  int* class_count;               // [ncls_cap] or nullptr
};

// bytes of shared memory nms_resolve needs (excluding class_count and the caller's kept list), 8-byte aligned
__host__ __device__ inline size_t nms_resolve_smem_bytes(int words) {
  return sizeof(unsigned long long) * (size_t)(words + NMS_SUPER_BOXES * NMS_SUPER + NMS_SUPER) +
         sizeof(float) * NMS_SUPER_BOXES + sizeof(int) * (NMS_SUPER_BOXES + 4);
}

__device__ __forceinline__ NmsResolveSmem nms_resolve_carve(unsigned long long* base, int words, int** rest) {
  NmsResolveSmem s;
  s.remv = base;
  s.diag = s.remv + words;
  s.svalid = s.diag + NMS_SUPER_BOXES * NMS_SUPER;
  s.scls = (float*)(s.svalid + NMS_SUPER);
  s.kept_rows = (int*)(s.scls + NMS_SUPER_BOXES);
  s.misc = s.kept_rows + NMS_SUPER_BOXES;
  s.class_count = nullptr;
  *rest = s.misc + 4;
  return s;
}

__device__ __forceinline__ int nms_resolve(const float4* __restrict__ boxes,
                                           const float* __restrict__ cls,
                                           const unsigned long long* __restrict__ mask,
                                           int n, int words, int max_total,
                                           int max_per_class, int ncls_cap,
                                           NmsResolveSmem s, int* kept_out) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nchunks = (n + NMS_TILE - 1) / NMS_TILE;               // mask words in use
  const int nsuper = (nchunks + NMS_SUPER - 1) / NMS_SUPER;
  for (int w = tid; w < words; w += nt) s.remv[w] = 0ull;
  if (s.class_count) for (int k = tid; k < ncls_cap; k += nt) s.class_count[k] = 0;
  if (tid == 0) { s.misc[0] = 0; s.misc[1] = 0; s.misc[2] = 0; }
  __syncthreads();
  unsigned int* svalid32 = reinterpret_cast<unsigned int*>(s.svalid);
  for (int c = 0; c < nsuper; ++c) {
    const int row0 = c * NMS_SUPER_BOXES;
    const int w0 = c * NMS_SUPER;
    // step 1: stage the intra-chunk mask block, selectability and classes of the 256 rows of this super-chunk
    if (tid < NMS_SUPER_BOXES) {
      const int row = row0 + tid;
      bool ok = false;
      float cv = 0.0f;
      const int myw = tid / NMS_TILE;                                // word (within the super-chunk) this row lives in
      #pragma unroll
      for (int j = 0; j < NMS_SUPER; ++j) {
        unsigned long long d = 0ull;
        if (row < n && j >= myw && w0 + j < nchunks) d = mask[(size_t)row * words + w0 + j];   // only words >= the row's own were written
        s.diag[tid * NMS_SUPER + j] = d;
      }
      if (row < n) { ok = box_selectable(boxes[row]); if (cls) cv = cls[row]; }
      s.scls[tid] = cv;
      const unsigned int bal = __ballot_sync(0xffffffffu, ok);
      if ((tid & 31) == 0) svalid32[tid >> 5] = bal;
    }
    __syncthreads();
    // step 2: one thread walks the alive candidates in order (shared memory only)
    if (tid == 0) {
      unsigned long long alive[NMS_SUPER];
      #pragma unroll
      for (int j = 0; j < NMS_SUPER; ++j) alive[j] = (w0 + j < nchunks) ? (s.svalid[j] & ~s.remv[w0 + j]) : 0ull;
      int total = s.misc[0];
      int nk = 0;
      #pragma unroll
      for (int j = 0; j < NMS_SUPER; ++j) {
        while (alive[j] && total < max_total) {
          const int t = __ffsll((long long)alive[j]) - 1;
          alive[j] &= ~(1ull << t);
          const int idx = j * NMS_TILE + t;
          if (s.class_count) {
            int k = (int)s.scls[idx];
            k = k < 0 ? 0 : (k >= ncls_cap ? ncls_cap - 1 : k);
            if (s.class_count[k] >= max_per_class) continue;     // Utils.swift:192 per class call
            s.class_count[k]++;
          }
          #pragma unroll
          for (int jj = 0; jj < NMS_SUPER; ++jj) if (jj >= j) alive[jj] &= ~s.diag[idx * NMS_SUPER + jj];
          kept_out[total] = row0 + idx;
          s.kept_rows[nk++] = row0 + idx;
          total++;
        }
      }
      s.misc[0] = total;
      s.misc[1] = nk;
      s.misc[2] = (total >= max_total) ? 1 : 0;
    }
    __syncthreads();
    const int nk = s.misc[1];
    const int stop = s.misc[2];
    if (stop) break;
    // step 3: OR the kept rows into the removal words of the later super-chunks.  (kept row, word) pairs are
    // flattened over all threads: consecutive threads read consecutive words of one mask row (coalesced), every
    // load is independent, and non-zero words are merged with shared-memory atomics.
    const int wfirst = w0 + NMS_SUPER;
    const int nlater = nchunks - wfirst;
    if (nk > 0 && nlater > 0) {
      const int pairs = nk * nlater;
      for (int p0 = tid; p0 < pairs; p0 += 4 * nt) {
        unsigned long long v[4];
        int wv[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int pp = p0 + u * nt;
          v[u] = 0ull; wv[u] = wfirst;
          if (pp < pairs) {
            const int k = pp / nlater;
            wv[u] = wfirst + (pp - k * nlater);
            v[u] = mask[(size_t)s.kept_rows[k] * words + wv[u]];
          }
        }
        #pragma unroll
        for (int u = 0; u < 4; ++u) if (v[u]) atomicOr(&s.remv[wv[u]], v[u]);
      }
    }
    __syncthreads();
  }
  __syncthreads();
  return s.misc[0];
}
