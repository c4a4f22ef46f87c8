// detection.cu -- DetectionLayer.evaluate (DetectionLayer.swift:107-234) and the
// small post-processing halves of TimeDistributedClassifierLayer (:50-88) and
// Detection.swift (:23-99).
//
// Per image:
//   det_filter_kernel   score >= thr (vDSP_vthres, :259) && classId > 0 (:136-140),
//                       ORDERED compaction (roi order is the NMS visiting order,
//                       Q10), gather, x std, decode, clip (:144-164)
//   nms_mask_kernel     same-class pairs only (per-class NMS, :170-183)
//   det_finalize_percls_kernel  per-class greedy resolve (one thread per class, bitmask rows staged in
//                       shared memory) with a per-class cap of maxDetections, then top maxDetections by
//                       (score desc, class asc, position asc) (:186-209, intended-mode ties), zero pad;
//                       det_finalize_kernel = the generic sequential resolve, used when the per-class
//                       bitmaps do not fit in shared memory (> ~1400 rois).
#include "common.cuh"
#include "nms.cuh"

#define DET_THREADS 1024
#define DET_MAX_CLASSES 1024

__global__ void __launch_bounds__(DET_THREADS)
det_filter_kernel(const float4* __restrict__ rois, const float* __restrict__ cls, int R,
                  float4 sd, float score_thr, float4* __restrict__ fbox, float* __restrict__ fcls,
                  float* __restrict__ fscore, int32_t* __restrict__ fidx, int32_t* __restrict__ fcount) {
  __shared__ int s_warp[DET_THREADS / 32];
  __shared__ int s_base, s_round;
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float4* r = rois + (size_t)img * R;
  const float* c = cls + (size_t)img * R * 6;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int start = 0; start < R; start += DET_THREADS) {
    int i = start + tid;
    bool keep = false;
    float score = 0.f, classId = 0.f;
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < R) {
      const float* row = c + (size_t)i * 6;
      d = make_float4(row[0], row[1], row[2], row[3]);
      classId = row[4]; score = row[5];
      keep = (score >= score_thr) && (score != 0.0f) && (classId > 0.0f);
    }
    unsigned int bal = __ballot_sync(0xffffffffu, keep);
    int wcount = __popc(bal);
    if (lane == 0) s_warp[wid] = wcount;
    __syncthreads();
    if (wid == 0) {
      int v = s_warp[lane];
      int iv = v;
      #pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += u; }
      s_warp[lane] = iv - v;              // exclusive prefix of the warp counts
      if (lane == 31) s_round = iv;       // number kept in this round
    }
    __syncthreads();
    int pos = s_base + s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
    if (keep) {
      float4 box = decode_box(r[i], d, sd);
      size_t o = (size_t)img * R + pos;
      fbox[o] = box; fcls[o] = classId; fscore[o] = score; fidx[o] = i;
    }
    __syncthreads();
    if (tid == 0) s_base += s_round;
    __syncthreads();
  }
  if (tid == 0) fcount[img] = s_base;
}

__global__ void __launch_bounds__(DET_THREADS)
det_finalize_kernel(const float4* __restrict__ fbox, const float* __restrict__ fcls,
                    const float* __restrict__ fscore, const int32_t* __restrict__ fidx,
                    const int32_t* __restrict__ fcount, const unsigned long long* __restrict__ mask,
                    int R, int words, int max_det, int ncls_cap, float* __restrict__ out,
                    int32_t* __restrict__ keep_roi, int32_t* __restrict__ count_out) {
  extern __shared__ unsigned long long smem_u64[];
  const int img = blockIdx.x, tid = threadIdx.x;
  const int K = fcount[img];
  int* rest = nullptr;
  NmsResolveSmem s = nms_resolve_carve(smem_u64, words, &rest);
  s.class_count = rest;
  int* kept = s.class_count + ncls_cap;   // [R]
  float* kscore = (float*)(kept + R);     // [R]
  float* kcls = kscore + R;               // [R]
  const float4* b = fbox + (size_t)img * R;
  const float* c = fcls + (size_t)img * R;
  const float* sc = fscore + (size_t)img * R;
  const unsigned long long* mk = mask + (size_t)img * R * words;
  int nk = nms_resolve(b, c, mk, K, words, 0x7fffffff, max_det, ncls_cap, s, kept);
  for (int i = tid; i < nk; i += DET_THREADS) { kscore[i] = sc[kept[i]]; kcls[i] = c[kept[i]]; }
  // zero the whole output first (rows >= count must be zero, :228-231)
  for (int i = tid; i < max_det * 6; i += DET_THREADS) out[(size_t)img * max_det * 6 + i] = 0.0f;
  if (keep_roi) for (int i = tid; i < max_det; i += DET_THREADS) keep_roi[(size_t)img * max_det + i] = -1;
  __syncthreads();
  // rank of every kept element in the order (score desc, class asc, position asc)
  for (int e = tid; e < nk; e += DET_THREADS) {
    float se = kscore[e], ce = kcls[e];
    int pe = kept[e];
    int rank = 0;
    for (int f = 0; f < nk; ++f) {
      float sf = kscore[f], cf = kcls[f];
      int pf = kept[f];
      bool before = (sf > se) || (sf == se && (cf < ce || (cf == ce && pf < pe)));
      rank += before ? 1 : 0;
    }
    if (rank < max_det) {
      float4 box = b[pe];
      float* o = out + ((size_t)img * max_det + rank) * 6;
      o[0] = box.x; o[1] = box.y; o[2] = box.z; o[3] = box.w; o[4] = ce; o[5] = se;
      if (keep_roi) keep_roi[(size_t)img * max_det + rank] = fidx[(size_t)img * R + pe];
    }
  }
  if (count_out && tid == 0) count_out[img] = nk < max_det ? nk : max_det;
}

// Per-class NMS is independent per class (the bitmask only links same-class pairs), so the order-dependent walk
// is done by one thread PER CLASS in parallel: class k's thread visits its candidates in roi order, keeps at most
// max_det of them and ORs the kept rows into its private removal words.  Then the kept set is compacted and ranked
// exactly like det_finalize_kernel does (the final order does not depend on the order of the kept list).
__global__ void __launch_bounds__(DET_THREADS)
det_finalize_percls_kernel(const float4* __restrict__ fbox, const float* __restrict__ fcls,
                           const float* __restrict__ fscore, const int32_t* __restrict__ fidx,
                           const int32_t* __restrict__ fcount, const unsigned long long* __restrict__ mask,
                           int R, int words, int max_det, int ncls_cap, float* __restrict__ out,
                           int32_t* __restrict__ keep_roi, int32_t* __restrict__ count_out) {
  extern __shared__ unsigned long long smem_u64[];
  const int img = blockIdx.x, tid = threadIdx.x;
  const int K = fcount[img];
  const int nwords = (K + 63) >> 6;
  unsigned long long* cbits = smem_u64;                              // [ncls][words] candidates of each class
  unsigned long long* rem = cbits + (size_t)ncls_cap * words;        // [ncls][words] suppressed by a kept box of that class
  unsigned long long* valid = rem + (size_t)ncls_cap * words;        // [words] selectable (Utils.swift:195)
  unsigned long long* keptbits = valid + words;                      // [words]
  int* kept = (int*)(keptbits + words);                              // [R]
  float* kscore = (float*)(kept + R);                                // [R]
  float* kcls = kscore + R;                                          // [R]
  int* counter = (int*)(kcls + R);
  unsigned long long* smask = (unsigned long long*)(((uintptr_t)(counter + 1) + 7) & ~(uintptr_t)7);   // [K][nwords] staged bitmask rows
  const float4* b = fbox + (size_t)img * R;
  const float* c = fcls + (size_t)img * R;
  const float* sc = fscore + (size_t)img * R;
  const unsigned long long* mk = mask + (size_t)img * R * words;
  for (int i = tid; i < 2 * ncls_cap * words + 2 * words; i += DET_THREADS) smem_u64[i] = 0ull;
  if (tid == 0) *counter = 0;
  // stage the (upper-triangular) bitmask rows of the K candidates: the per-class walks then touch shared memory only
  for (int e = tid; e < K * nwords; e += DET_THREADS) {
    const int i = e / nwords, w = e - i * nwords;
    smask[e] = (w >= (i >> 6)) ? mk[(size_t)i * words + w] : 0ull;
  }
  __syncthreads();
  for (int i = tid; i < K; i += DET_THREADS) {
    int k = (int)c[i];
    k = k < 0 ? 0 : (k >= ncls_cap ? ncls_cap - 1 : k);
    const unsigned long long bit = 1ull << (i & 63);
    atomicOr(&cbits[(size_t)k * words + (i >> 6)], bit);
    if (box_selectable(b[i])) atomicOr(&valid[i >> 6], bit);
  }
  __syncthreads();
  if (tid < ncls_cap && nwords <= 16) {
    // removal words of this class live in registers (R <= 1024): the row ORs of a kept box are independent loads
    const unsigned long long* cb = cbits + (size_t)tid * words;
    unsigned long long rm[16];
    #pragma unroll
    for (int ww = 0; ww < 16; ++ww) rm[ww] = 0ull;
    int count = 0;
    #pragma unroll
    for (int w = 0; w < 16; ++w) {
      if (w < nwords && count < max_det) {
        unsigned long long alive = cb[w] & valid[w] & ~rm[w];
        while (alive && count < max_det) {
          const int t = __ffsll((long long)alive) - 1;
          const unsigned long long bit = 1ull << t;
          alive &= ~bit;
          const int i = (w << 6) + t;
          atomicOr(&keptbits[w], bit);
          ++count;
          const unsigned long long* row = smask + (size_t)i * nwords;
          alive &= ~row[w];
          #pragma unroll
          for (int ww = 0; ww < 16; ++ww) if (ww > w && ww < nwords) rm[ww] |= row[ww];
        }
      }
    }
  } else if (tid < ncls_cap) {
    const unsigned long long* cb = cbits + (size_t)tid * words;
    unsigned long long* rm = rem + (size_t)tid * words;
    int count = 0;
    for (int w = 0; w < nwords && count < max_det; ++w) {
      unsigned long long alive = cb[w] & valid[w] & ~rm[w];
      while (alive && count < max_det) {                  // Utils.swift:192: at most max per class call
        const int t = __ffsll((long long)alive) - 1;
        const unsigned long long bit = 1ull << t;
        alive &= ~bit;
        const int i = (w << 6) + t;
        atomicOr(&keptbits[w], bit);
        ++count;
        const unsigned long long* row = smask + (size_t)i * nwords;
        alive &= ~row[w];
        for (int ww = w + 1; ww < nwords; ++ww) rm[ww] |= row[ww];
      }
    }
  }
  // zero the whole output first (rows >= count must be zero, DetectionLayer.swift:228-231)
  for (int i = tid; i < max_det * 6; i += DET_THREADS) out[(size_t)img * max_det * 6 + i] = 0.0f;
  if (keep_roi) for (int i = tid; i < max_det; i += DET_THREADS) keep_roi[(size_t)img * max_det + i] = -1;
  __syncthreads();
  for (int i = tid; i < K; i += DET_THREADS)
    if ((keptbits[i >> 6] >> (i & 63)) & 1ull) {
      const int slot = atomicAdd(counter, 1);
      kept[slot] = i; kscore[slot] = sc[i]; kcls[slot] = c[i];
    }
  __syncthreads();
  const int nk = *counter;
  // rank of every kept element in the order (score desc, class asc, position asc)
  for (int e = tid; e < nk; e += DET_THREADS) {
    const float se = kscore[e], ce = kcls[e];
    const int pe = kept[e];
    int rank = 0;
    for (int f = 0; f < nk; ++f) {
      const float sf = kscore[f], cf = kcls[f];
      const int pf = kept[f];
      rank += ((sf > se) || (sf == se && (cf < ce || (cf == ce && pf < pe)))) ? 1 : 0;
    }
    if (rank < max_det) {
      const float4 box = b[pe];
      float* o = out + ((size_t)img * max_det + rank) * 6;
      o[0] = box.x; o[1] = box.y; o[2] = box.z; o[3] = box.w; o[4] = ce; o[5] = se;
      if (keep_roi) keep_roi[(size_t)img * max_det + rank] = fidx[(size_t)img * R + pe];
    }
  }
  if (count_out && tid == 0) count_out[img] = nk < max_det ? nk : max_det;
}

static int detection_ensure_ws(mrcnn_ctx* ctx, int batch, int64_t R) {
  if (batch <= ctx->det_batch && R <= ctx->det_rois) return MRCNN_OK;
  MRCNN_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->d_fbox); cudaFree(ctx->d_fcls); cudaFree(ctx->d_fscore);
  cudaFree(ctx->d_fidx); cudaFree(ctx->d_fcount); cudaFree(ctx->d_dmask);
  int b = batch > ctx->det_batch ? batch : ctx->det_batch;
  int64_t r = R > ctx->det_rois ? R : ctx->det_rois;
  int words = ceil_div(r, 64);
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_fbox, sizeof(float4) * r * b));
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_fcls, sizeof(float) * r * b));
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_fscore, sizeof(float) * r * b));
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_fidx, sizeof(int32_t) * r * b));
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_fcount, sizeof(int32_t) * b));
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_dmask, sizeof(unsigned long long) * r * words * b));
  ctx->det_batch = b; ctx->det_rois = r;
  return MRCNN_OK;
}

int detection_run(mrcnn_ctx* ctx, int batch, int64_t R64, const float* d_rois, const float* d_cls,
                  float* d_out, int32_t* d_keep_roi, int32_t* d_count) {
  const mrcnn_config& cfg = ctx->cfg;
  MRCNN_REQUIRE(ctx, batch >= 1, "detection: batch must be >= 1");
  MRCNN_REQUIRE(ctx, R64 >= 1 && R64 <= 8192, "detection: num_rois must be in [1, 8192]");
  MRCNN_REQUIRE(ctx, cfg.max_detections >= 1 && cfg.max_detections <= 4096, "detection: max_detections out of range");
  MRCNN_REQUIRE(ctx, cfg.num_classes >= 1 && cfg.num_classes <= DET_MAX_CLASSES, "detection: num_classes out of range");
  int rc = detection_ensure_ws(ctx, batch, R64);
  if (rc) return rc;
  // workspace arrays are indexed with stride R (the allocation holds det_rois >= R rois per image)
  const int R = (int)R64;
  const int stride = R;
  const int words = ceil_div(R, 64);
  cudaStream_t s = ctx->stream;
  float4 sd = make_float4(cfg.bbox_std[0], cfg.bbox_std[1], cfg.bbox_std[2], cfg.bbox_std[3]);
  ProfScope ps(ctx, PROF_DET_FILTER, (double)batch * R * (16 + 24 + 28));   // whole layer: filter + mask + finalize
  det_filter_kernel<<<batch, DET_THREADS, 0, s>>>((const float4*)d_rois, d_cls, R, sd, cfg.detection_min_score,
                                                   ctx->d_fbox, ctx->d_fcls, ctx->d_fscore, ctx->d_fidx, ctx->d_fcount);
  MRCNN_LAUNCH_CHECK(ctx);
  const int tiles = ceil_div(R, NMS_TILE);
  dim3 mgrid(tiles, tiles, batch);
  nms_mask_kernel<<<mgrid, NMS_TILE, 0, s>>>(ctx->d_fbox, ctx->d_fcls, ctx->d_fcount, 0, stride, words,
                                             cfg.detection_nms_iou, ctx->d_dmask);
  MRCNN_LAUNCH_CHECK(ctx);
  const int ncls_cap = cfg.num_classes;
  // per-class parallel resolve when its bitmaps fit in shared memory, else the generic sequential resolve
  const size_t sm_pc = sizeof(unsigned long long) * (2 * (size_t)ncls_cap * words + 2 * words) + sizeof(int) * (R + 1) + sizeof(float) * 2 * R +
                       16 + sizeof(unsigned long long) * (size_t)R * words;
  if (sm_pc <= 200 * 1024 && ncls_cap <= DET_THREADS) {
    if (sm_pc > 48 * 1024)
      MRCNN_CUDA_TRY(ctx, cudaFuncSetAttribute(det_finalize_percls_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_pc));
    det_finalize_percls_kernel<<<batch, DET_THREADS, sm_pc, s>>>(ctx->d_fbox, ctx->d_fcls, ctx->d_fscore, ctx->d_fidx,
                                                                 ctx->d_fcount, ctx->d_dmask, R, words, cfg.max_detections,
                                                                 ncls_cap, d_out, d_keep_roi, d_count);
  } else {
    size_t sm = nms_resolve_smem_bytes(words) + sizeof(int) * (ncls_cap + R) + sizeof(float) * 2 * R;
    MRCNN_REQUIRE(ctx, sm <= 200 * 1024, "detection: workspace exceeds shared memory");
    if (sm > 48 * 1024) {
      MRCNN_CUDA_TRY(ctx, cudaFuncSetAttribute(det_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    }
    det_finalize_kernel<<<batch, DET_THREADS, sm, s>>>(ctx->d_fbox, ctx->d_fcls, ctx->d_fscore, ctx->d_fidx,
                                                       ctx->d_fcount, ctx->d_dmask, R, words, cfg.max_detections,
                                                       ncls_cap, d_out, d_keep_roi, d_count);
  }
  MRCNN_LAUNCH_CHECK(ctx);
  return MRCNN_OK;
}

// ---------------------------------------------------------------------------
// TimeDistributedClassifierLayer.swift:50-88 + :177-192 (vDSP_maxvi: first max)
// One warp per roi: probabilities (ncls), bounding_boxes (ncls,4) -> (6).
// ---------------------------------------------------------------------------
__global__ void classifier_select_kernel(const float* __restrict__ probs, const float* __restrict__ bbox,
                                         int64_t total, int ncls, float* __restrict__ out) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= total) return;
  const float* p = probs + r * ncls;
  float bv = -INFINITY; int bi = 0x7fffffff;
  for (int c = lane; c < ncls; c += 32) {
    float v = p[c];
    if (v > bv) { bv = v; bi = c; }       // strict '>' keeps the first max per lane
  }
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (bi == 0x7fffffff) { bi = 0; bv = p[0]; }   // all NaN / -inf: index 0 like vDSP_maxvi's init
  if (lane < 4) out[r * 6 + lane] = bbox[(r * ncls + bi) * 4 + lane];
  if (lane == 4) out[r * 6 + 4] = (float)bi;
  if (lane == 5) out[r * 6 + 5] = bv;
}

int classifier_select_run(mrcnn_ctx* ctx, int batch, int64_t R, int ncls, const float* d_probs,
                          const float* d_bbox, float* d_out) {
  int64_t total = (int64_t)batch * R;
  int threads = 256;
  int blocks = ceil_div(total * 32, threads);
  ProfScope ps(ctx, PROF_GLUE, (double)total * (ncls * 4.0 + 16 + 24));
  classifier_select_kernel<<<blocks, threads, 0, ctx->stream>>>(d_probs, d_bbox, total, ncls, d_out);
  MRCNN_LAUNCH_CHECK(ctx);
  return MRCNN_OK;
}

// ---------------------------------------------------------------------------
// Detection.swift:23-62 (score > 0.7 in Double, bbox x,y,w,h) and :64-99 (8-bit
// mask 255 - p/2*255, truncation).  One CTA per image.
// ---------------------------------------------------------------------------
__global__ void detections_decode_kernel(const float* __restrict__ det, const float* __restrict__ masks,
                                         int D, int S, int32_t* __restrict__ count_out,
                                         int32_t* __restrict__ index_out, double* __restrict__ bbox_out,
                                         int32_t* __restrict__ class_out, double* __restrict__ score_out,
                                         uint8_t* __restrict__ mask_out) {
  extern __shared__ int s_slot[];   // [D] output slot per detection or -1
  const int img = blockIdx.x, tid = threadIdx.x;
  const float* d = det + (size_t)img * D * 6;
  if (tid == 0) {
    int n = 0;
    for (int i = 0; i < D; ++i) {
      double score = (double)d[i * 6 + 5];
      s_slot[i] = (score > 0.7) ? n++ : -1;          // Detection.swift:38
    }
    count_out[img] = n;
  }
  __syncthreads();
  const int plane = S * S;
  for (int i = tid; i < D; i += blockDim.x) {
    // zero row i (rows >= count stay zero)
    index_out[(size_t)img * D + i] = 0; class_out[(size_t)img * D + i] = 0; score_out[(size_t)img * D + i] = 0.0;
    for (int k = 0; k < 4; ++k) bbox_out[((size_t)img * D + i) * 4 + k] = 0.0;
  }
  if (mask_out) for (int i = tid; i < D * plane; i += blockDim.x) mask_out[(size_t)img * D * plane + i] = 0;
  __syncthreads();
  for (int i = tid; i < D; i += blockDim.x) {
    int n = s_slot[i];
    if (n < 0) continue;
    double y1 = d[i * 6], x1 = d[i * 6 + 1], y2 = d[i * 6 + 2], x2 = d[i * 6 + 3];
    size_t o = (size_t)img * D + n;
    index_out[o] = i;
    bbox_out[o * 4] = x1; bbox_out[o * 4 + 1] = y1;
    bbox_out[o * 4 + 2] = __dsub_rn(x2, x1); bbox_out[o * 4 + 3] = __dsub_rn(y2, y1);
    class_out[o] = (int)d[i * 6 + 4];
    score_out[o] = (double)d[i * 6 + 5];
  }
  if (mask_out && masks) {
    for (int e = tid; e < D * plane; e += blockDim.x) {
      int i = e / plane, p = e - i * plane;
      int n = s_slot[i];
      if (n < 0) continue;
      double v = (double)masks[((size_t)img * D + i) * plane + p];
      double bq = __dsub_rn(255.0, __dmul_rn(__ddiv_rn(v, 2.0), 255.0));   // :84
      bq = bq < 0.0 ? 0.0 : (bq > 255.0 ? 255.0 : bq);   // Swift would trap outside UInt8; sigmoid keeps it in range
      mask_out[((size_t)img * D + n) * plane + p] = (uint8_t)bq;
    }
  }
}

int detections_decode_run(mrcnn_ctx* ctx, int batch, int D, int S, const float* d_det, const float* d_masks,
                          int32_t* d_count, int32_t* d_index, double* d_bbox, int32_t* d_class,
                          double* d_score, uint8_t* d_mask_u8) {
  ProfScope ps(ctx, PROF_GLUE, (double)batch * D * (24.0 + (d_masks ? 5.0 * S * S : 0.0) + 56.0));
  detections_decode_kernel<<<batch, 256, sizeof(int) * D, ctx->stream>>>(d_det, d_masks, D, S, d_count, d_index,
                                                                         d_bbox, d_class, d_score, d_mask_u8);
  MRCNN_LAUNCH_CHECK(ctx);
  return MRCNN_OK;
}
