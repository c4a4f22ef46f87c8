// common.cuh -- context, error handling and launch helpers shared by every TU of
// libmaskrcnn_cuda.so.  sm_100a only; there is no CPU fallback anywhere.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>

#include "../../include/maskrcnn_cuda.h"

#define MRCNN_SM_COUNT_FALLBACK 148

struct DenseModel;  // dense.cu (backbone / heads), opaque here

struct mrcnn_ctx {
  mrcnn_config cfg;
  std::string anchors_path, main_path, cls_path, mask_path;
  int device = 0;
  int sm_count = MRCNN_SM_COUNT_FALLBACK;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaMemPool_t pool = nullptr;       // staging pool for host-pointer calls (keeps its memory between calls)
  std::string err;
  int64_t launches = 0;

  // anchors (N,4) f32 on device
  float* d_anchors = nullptr;
  int64_t num_anchors = 0;

  // ---- proposal workspace (sized for ws_batch x ws_anchors) ----
  int ws_batch = 0;
  int64_t ws_anchors = 0;
  int ws_pre = 0;
  uint32_t* d_hist = nullptr;        // [B][RADIX_BINS]
  struct SelState* d_sel = nullptr;  // [B]
  unsigned long long* d_cand = nullptr;  // [B][sort_n]
  float4* d_sboxes = nullptr;        // [B][pre] decoded boxes in score order
  int32_t* d_sorder = nullptr;       // [B][pre] anchor index in score order
  unsigned long long* d_mask = nullptr;  // [B][pre][words]
  size_t mask_bytes = 0;

  // ---- detection workspace (sized for det_batch x det_rois) ----
  int det_batch = 0;
  int64_t det_rois = 0;
  float4* d_fbox = nullptr;   // [B][R]
  float* d_fcls = nullptr;    // [B][R]
  float* d_fscore = nullptr;  // [B][R]
  int32_t* d_fidx = nullptr;  // [B][R]
  int32_t* d_fcount = nullptr;  // [B]
  unsigned long long* d_dmask = nullptr;  // [B][R][words]

  // ---- roialign workspace ----
  int roi_cap = 0;            // batch*R capacity
  int32_t* d_roi_level = nullptr;
  void* roi_tma = nullptr;    // roialign.cu: tensor maps of the staged kernel (RoiTmaCache)

  // ---- dense model (backbone / heads) ----
  DenseModel* dense = nullptr;

  // ---- NCCL ----
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  float* d_gather_send = nullptr;
  float* d_gather_recv = nullptr;
  size_t gather_bytes = 0;

  // stage timing of last predict
  std::vector<std::pair<const char*, float>> stage_ms;

  // ---- per-kernel-class device timing (mrcnn_profile_*) ----
  bool profiling = false;
  struct ProfRec { int cls; cudaEvent_t e0, e1; double work; int launches; };
  std::vector<ProfRec> prof_recs;      // recorded since the last read
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pool;  // reusable event pairs
  size_t prof_used = 0;
};

// Kernel classes reported by mrcnn_profile_read (names in api.cu, same order).
enum ProfClass {
  PROF_CONV_GEMM = 0, PROF_ROIALIGN, PROF_TOPK_SELECT, PROF_SORT_DECODE, PROF_NMS_MASK, PROF_NMS_RESOLVE,
  PROF_DET_FILTER, PROF_DET_FINALIZE, PROF_GLUE, PROF_NUM_CLASSES
};

// RAII: brackets the launches issued in its scope with a CUDA event pair on the
// context's stream when profiling is on (no-op otherwise).  `work` = algorithmic
// bytes (memory-bound classes) or flops (PROF_CONV_GEMM) of those launches.
struct ProfScope {
  mrcnn_ctx* ctx; int idx = -1;
  ProfScope(mrcnn_ctx* c, int cls, double work) : ctx(c) {
    if (!c->profiling) return;
    // consecutive scopes of the same class share one event pair (only the end event is re-recorded), so that
    // back-to-back launches of a class are timed as they run in production: no event between them
    if (!c->prof_recs.empty() && c->prof_recs.back().cls == cls) {
      idx = (int)c->prof_recs.size() - 1;
      c->prof_recs[idx].work += work;
      c->prof_recs[idx].launches += 1;
      return;
    }
    if (c->prof_used == c->prof_pool.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
      c->prof_pool.push_back({a, b});
    }
    auto& pr = c->prof_pool[c->prof_used++];
    cudaEventRecord(pr.first, c->stream);
    c->prof_recs.push_back({cls, pr.first, pr.second, work, 1});
    idx = (int)c->prof_recs.size() - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(ctx->prof_recs[idx].e1, ctx->stream); }
};

inline int mrcnn_fail(mrcnn_ctx* ctx, int code, const std::string& msg);

#define MRCNN_CUDA_TRY(ctx, expr)                                              \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      char _b[512];                                                            \
      snprintf(_b, sizeof(_b), "%s:%d: %s failed: %s", __FILE__, __LINE__,    \
               #expr, cudaGetErrorString(_e));                                 \
      return mrcnn_fail((ctx), MRCNN_ECUDA, _b);                               \
    }                                                                          \
  } while (0)

#define MRCNN_REQUIRE(ctx, cond, msg)                                          \
  do {                                                                         \
    if (!(cond)) return mrcnn_fail((ctx), MRCNN_EINVAL, std::string(msg));     \
  } while (0)

#define MRCNN_LAUNCH_CHECK(ctx)                                                \
  do {                                                                         \
    (ctx)->launches++;                                                         \
    cudaError_t _e = cudaGetLastError();                                       \
    if (_e != cudaSuccess) {                                                   \
      char _b[512];                                                            \
      snprintf(_b, sizeof(_b), "%s:%d: kernel launch failed: %s", __FILE__,   \
               __LINE__, cudaGetErrorString(_e));                              \
      return mrcnn_fail((ctx), MRCNN_ECUDA, _b);                               \
    }                                                                          \
  } while (0)

std::string& mrcnn_tls_error();

inline int mrcnn_fail(mrcnn_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  mrcnn_tls_error() = msg;
  return code;
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------
// Host/device pointer staging for the layer-level ABI (host pointers allowed).
// Device pointers pass straight through; host pointers are mirrored in a
// stream-ordered allocation (pool-backed, no cudaMalloc after warm-up).
// ---------------------------------------------------------------------------
struct Staged {
  void* dev = nullptr;
  void* host = nullptr;  // non-null iff the caller's pointer was a host pointer
  size_t bytes = 0;
  bool is_output = false;
};

struct Stager {
  mrcnn_ctx* ctx;
  std::vector<Staged> items;
  bool any_host = false;
  explicit Stager(mrcnn_ctx* c) : ctx(c) {}
  static bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
  }
  // returns device pointer (nullptr on failure with *st set)
  void* in(const void* p, size_t bytes, int* st) {
    if (!p || bytes == 0) return const_cast<void*>(p);
    if (is_device_ptr(p)) return const_cast<void*>(p);
    Staged s; s.host = const_cast<void*>(p); s.bytes = bytes; s.is_output = false;
    if (cudaMallocFromPoolAsync(&s.dev, bytes, ctx->pool, ctx->stream) != cudaSuccess) { *st = MRCNN_ECUDA; return nullptr; }
    if (cudaMemcpyAsync(s.dev, p, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { *st = MRCNN_ECUDA; return nullptr; }
    items.push_back(s); any_host = true;
    return s.dev;
  }
  void* out(void* p, size_t bytes, int* st) {
    if (!p || bytes == 0) return p;
    if (is_device_ptr(p)) return p;
    Staged s; s.host = p; s.bytes = bytes; s.is_output = true;
    if (cudaMallocFromPoolAsync(&s.dev, bytes, ctx->pool, ctx->stream) != cudaSuccess) { *st = MRCNN_ECUDA; return nullptr; }
    items.push_back(s); any_host = true;
    return s.dev;
  }
  // copies outputs back, frees staging, synchronises if anything was host-side
  int finish() {
    int rc = MRCNN_OK;
    for (auto& s : items)
      if (s.is_output)
        if (cudaMemcpyAsync(s.host, s.dev, s.bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = MRCNN_ECUDA;
    for (auto& s : items) cudaFreeAsync(s.dev, ctx->stream);
    items.clear();
    if (any_host) {
      cudaError_t e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) return mrcnn_fail(ctx, MRCNN_ECUDA, std::string("stream sync: ") + cudaGetErrorString(e));
    }
    if (rc != MRCNN_OK) return mrcnn_fail(ctx, rc, "staging copy failed");
    return rc;
  }
  ~Stager() { for (auto& s : items) cudaFreeAsync(s.dev, ctx->stream); }
};

// ---- entry points implemented per TU (internal C++ linkage) -------------------
int proposal_run(mrcnn_ctx* ctx, int batch, int64_t N, const float* d_probs, const float* d_deltas,
                 float* d_rois_out, int32_t* d_keep_anchor, int32_t* d_count);
int detection_run(mrcnn_ctx* ctx, int batch, int64_t R, const float* d_rois, const float* d_cls,
                  float* d_out, int32_t* d_keep_roi, int32_t* d_count);
int roialign_chw_run(mrcnn_ctx* ctx, int batch, const float* d_rois, int roi_stride, int64_t R,
                     const float* const d_fmaps[4], const int32_t hw[8], int64_t C, int P,
                     float* d_out, int32_t* d_level_out);
int roialign_nhwc_f16_run(mrcnn_ctx* ctx, int batch, const float* d_rois, int roi_stride, int64_t R,
                          const __half* const d_fmaps[4], const int32_t hw[8], int64_t C, int P,
                          __half* d_out, int32_t* d_level_out);
void roialign_release(mrcnn_ctx* ctx);
int classifier_select_run(mrcnn_ctx* ctx, int batch, int64_t R, int ncls, const float* d_probs,
                          const float* d_bbox, float* d_out);
int detections_decode_run(mrcnn_ctx* ctx, int batch, int D, int S, const float* d_det, const float* d_masks,
                          int32_t* d_count, int32_t* d_index, double* d_bbox, int32_t* d_class,
                          double* d_score, uint8_t* d_mask_u8);
