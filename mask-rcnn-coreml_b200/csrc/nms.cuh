// nms.cuh -- greedy NMS (Utils.swift:185-218) as two device stages:
//   1. nms_mask_kernel: upper-triangular 64x64-tile suppression bitmask, IoU in
//      fp64 rounded to fp32 and compared with '>' (Utils.swift:203,232-246);
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
  __shared__ float col_cls[NMS_TILE];
  const int t = threadIdx.x;
  const int cj = cb * NMS_TILE + t;
  if (cj < n) {
    col_rect[t] = make_rect(b[cj]);
    col_cls[t] = c ? c[cj] : 0.0f;
  }
  __syncthreads();
  const int i = rb * NMS_TILE + t;
  if (i >= n) return;
  const RectD me = make_rect(b[i]);
  const float mycls = c ? c[i] : 0.0f;
  const int ncol = min(NMS_TILE, n - cb * NMS_TILE);
  unsigned long long bits = 0ull;
  const int j0 = (cb == rb) ? t + 1 : 0;
  for (int j = j0; j < ncol; ++j) {
    if (c && col_cls[j] != mycls) continue;
    if (iou_rect(col_rect[j], me) > thr) bits |= (1ull << j);
  }
  mask[((size_t)img * stride_boxes + i) * words + cb] = bits;
}

// Sequential resolution, executed by one CTA (blockDim.x >= 64, multiple of 32).
// shared scratch supplied by the caller:
//   remv[words], diag[64], kept_rows[64], s_misc[4], class_count[ncls_cap] (if cls)
// Returns (in every thread) the number of kept boxes; kept indices are appended
// to kept_out (shared or global) in visiting order.
struct NmsResolveSmem {
  unsigned long long* remv;       // [words]
  unsigned long long* diag;       // [64]
  int* kept_rows;                 // [64]
  int* misc;                      // [4]: 0 kept_total, 1 kept_in_chunk, 2 stop flag
  int* class_count;               // [ncls_cap] or nullptr
};

__device__ __forceinline__ int nms_resolve(const float4* __restrict__ boxes,
                                           const float* __restrict__ cls,
                                           const unsigned long long* __restrict__ mask,
                                           int n, int words, int max_total,
                                           int max_per_class, int ncls_cap,
                                           NmsResolveSmem s, int* kept_out) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nchunks = (n + NMS_TILE - 1) / NMS_TILE;
  for (int w = tid; w < words; w += nt) s.remv[w] = 0ull;
  if (s.class_count) for (int k = tid; k < ncls_cap; k += nt) s.class_count[k] = 0;
  if (tid == 0) { s.misc[0] = 0; s.misc[1] = 0; s.misc[2] = 0; }
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int row0 = c * NMS_TILE;
    // step 1: diagonal words + selectability of the 64 rows of this chunk
    unsigned long long my_valid = 0ull;
    if (tid < NMS_TILE) {
      int row = row0 + tid;
      unsigned long long d = 0ull;
      bool ok = false;
      if (row < n) {
        d = mask[(size_t)row * words + c];
        ok = box_selectable(boxes[row]);
      }
      s.diag[tid] = d;
      my_valid = ok ? 1ull : 0ull;
    }
    // ballot the valid bits of threads 0..63 (two warps) into one word
    unsigned int bal = __ballot_sync(0xffffffffu, my_valid != 0ull);
    if (tid == 0) s.kept_rows[0] = (int)bal;       // reuse as scratch (lo)
    if (tid == 32) s.kept_rows[1] = (int)bal;      // (hi)
    __syncthreads();
    // step 2: one thread walks the alive candidates of the chunk in order
    if (tid == 0) {
      unsigned long long valid = ((unsigned long long)(unsigned int)s.kept_rows[1] << 32) |
                                 (unsigned long long)(unsigned int)s.kept_rows[0];
      unsigned long long alive = valid & ~s.remv[c];
      int total = s.misc[0];
      int nk = 0;
      while (alive && total < max_total) {
        int t = __ffsll((long long)alive) - 1;
        alive &= ~(1ull << t);
        int row = row0 + t;
        if (s.class_count) {
          int k = (int)cls[row];
          k = k < 0 ? 0 : (k >= ncls_cap ? ncls_cap - 1 : k);
          if (s.class_count[k] >= max_per_class) continue;   // Utils.swift:192 per class call
          s.class_count[k]++;
        }
        alive &= ~s.diag[t];
        kept_out[total] = row;
        s.kept_rows[nk++] = row;
        total++;
      }
      s.misc[0] = total;
      s.misc[1] = nk;
      s.misc[2] = (total >= max_total) ? 1 : 0;
    }
    __syncthreads();
    const int nk = s.misc[1];
    const int stop = s.misc[2];
    if (stop) break;
    // step 3: OR the kept rows into the removal words of later chunks
    if (nk > 0) {
      for (int w = c + 1 + tid; w < nchunks; w += nt) {  // words >= nchunks; later words are never visited
        unsigned long long acc = 0ull;
        #pragma unroll 4
        for (int k = 0; k < nk; ++k) acc |= mask[(size_t)s.kept_rows[k] * words + w];
        s.remv[w] |= acc;
      }
    }
    __syncthreads();
  }
  __syncthreads();
  return s.misc[0];
}
