// api.cu -- the C ABI of libmaskrcnn_cuda.so (include/maskrcnn_cuda.h): context
// lifecycle and the layer-level entry points.  The dense model (backbone, heads,
// pipeline) lives in dense.cu / pipeline.cu.
#include "common.cuh"
#include "dense.h"
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <fstream>

std::string& mrcnn_tls_error() {
  static thread_local std::string e;
  return e;
}

extern "C" {

void mrcnn_config_default(mrcnn_config* c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  c->struct_size = (int32_t)sizeof(mrcnn_config);
  c->device = -1;
  c->image_h = 1024; c->image_w = 1024;           // PyramidROIAlignLayer.swift:46
  c->architecture = 101;                           // README.md:87
  c->num_classes = 81;                             // README.md:89
  c->bbox_std[0] = 0.1f; c->bbox_std[1] = 0.1f;    // ProposalLayer.swift:57
  c->bbox_std[2] = 0.2f; c->bbox_std[3] = 0.2f;
  c->pre_nms_max_proposals = 6000;                 // ProposalLayer.swift:59
  c->max_proposals = 1000;                         // ProposalLayer.swift:61
  c->proposal_nms_iou = 0.7f;                      // ProposalLayer.swift:63
  c->pool_size_classifier = 7;                     // PyramidROIAlignLayer.swift:45
  c->pool_size_mask = 14;
  c->fpn_selection_factor = 224.0f;                // PyramidROIAlignLayer.swift:98
  c->max_detections = 100;                         // DetectionLayer.swift:57
  c->detection_min_score = 0.7f;                   // DetectionLayer.swift:59
  c->detection_nms_iou = 0.3f;                     // DetectionLayer.swift:61
  c->mean_rgb[0] = 123.7f; c->mean_rgb[1] = 116.8f; c->mean_rgb[2] = 103.9f;  // Conversion/task.py:73-75
  c->max_batch = 8;
  c->precise_masks = 1;      // masks within 1e-4 of an fp32 evaluation (the tolerance of the path); 0 = faster, 5e-4
}

const char* mrcnn_version(void) { return "maskrcnn_cuda 0.2 (sm_100a, abi 2)"; }

const char* mrcnn_last_error(const mrcnn_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  return mrcnn_tls_error().c_str();
}

static int read_file(const std::string& path, std::vector<char>& out) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return -1;
  std::streamsize n = f.tellg();
  f.seekg(0);
  out.resize((size_t)n);
  if (n > 0 && !f.read(out.data(), n)) return -1;
  return 0;
}

int mrcnn_create(const mrcnn_config* cfg, mrcnn_ctx** out_ctx) {
  if (!cfg || !out_ctx) return mrcnn_fail(nullptr, MRCNN_EINVAL, "mrcnn_create: null argument");
  if (cfg->struct_size != (int32_t)sizeof(mrcnn_config))
    return mrcnn_fail(nullptr, MRCNN_EINVAL, "mrcnn_create: struct_size mismatch (use mrcnn_config_default)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return mrcnn_fail(nullptr, MRCNN_ECUDA, "mrcnn_create: no CUDA device (this library has no CPU fallback)");
  int dev = cfg->device;
  if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
  if (dev >= ndev) return mrcnn_fail(nullptr, MRCNN_EINVAL, "mrcnn_create: device ordinal out of range");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
    return mrcnn_fail(nullptr, MRCNN_ECUDA, "mrcnn_create: cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return mrcnn_fail(nullptr, MRCNN_ECUDA, "mrcnn_create: device is not sm_100 (Blackwell B200); kernels are built for sm_100a only");
  if (cudaSetDevice(dev) != cudaSuccess) return mrcnn_fail(nullptr, MRCNN_ECUDA, "mrcnn_create: cudaSetDevice failed");

  mrcnn_ctx* ctx = new mrcnn_ctx();
  ctx->cfg = *cfg;
  ctx->device = dev;
  ctx->sm_count = prop.multiProcessorCount;
  if (cfg->anchors_path) ctx->anchors_path = cfg->anchors_path;
  if (cfg->main_model_path) ctx->main_path = cfg->main_model_path;
  if (cfg->classifier_model_path) ctx->cls_path = cfg->classifier_model_path;
  if (cfg->mask_model_path) ctx->mask_path = cfg->mask_model_path;
  ctx->cfg.anchors_path = ctx->cfg.main_model_path = ctx->cfg.classifier_model_path = ctx->cfg.mask_model_path = nullptr;
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return mrcnn_fail(nullptr, MRCNN_ECUDA, "mrcnn_create: cudaStreamCreate failed");
  }
  ctx->stream = ctx->own_stream;
  {
    // stream-ordered staging pool that never gives memory back between calls (the default pool releases at every sync)
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    uint64_t keep = UINT64_MAX;
    if (cudaMemPoolCreate(&ctx->pool, &props) != cudaSuccess ||
        cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) {
      cudaStreamDestroy(ctx->own_stream);
      delete ctx;
      return mrcnn_fail(nullptr, MRCNN_ECUDA, "mrcnn_create: cudaMemPoolCreate failed");
    }
  }
  int rc = MRCNN_OK;
  if (!ctx->anchors_path.empty()) {
    std::vector<char> buf;
    if (read_file(ctx->anchors_path, buf) != 0 || buf.size() % 16 != 0 || buf.empty()) {
      rc = mrcnn_fail(nullptr, MRCNN_EIO, "mrcnn_create: cannot read anchors file " + ctx->anchors_path);
    } else {
      rc = mrcnn_set_anchors(ctx, (const float*)buf.data(), (int64_t)(buf.size() / 16));
    }
  }
  const std::string* paths[3] = {&ctx->main_path, &ctx->cls_path, &ctx->mask_path};
  for (int w = 0; w < 3 && rc == MRCNN_OK; ++w) {
    if (paths[w]->empty()) continue;
    std::vector<char> buf;
    if (read_file(*paths[w], buf) != 0) { rc = mrcnn_fail(nullptr, MRCNN_EIO, "mrcnn_create: cannot read weights file " + *paths[w]); break; }
    rc = mrcnn_set_weights(ctx, w, buf.data(), buf.size());
  }
  if (rc != MRCNN_OK) {
    std::string keep = ctx->err.empty() ? mrcnn_tls_error() : ctx->err;
    mrcnn_destroy(ctx);
    mrcnn_tls_error() = keep;
    return rc;
  }
  *out_ctx = ctx;
  return MRCNN_OK;
}

void mrcnn_destroy(mrcnn_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  dense_destroy(ctx);
  comm_destroy(ctx);
  cudaFree(ctx->d_anchors);
  cudaFree(ctx->d_hist); cudaFree(ctx->d_sel); cudaFree(ctx->d_cand);
  cudaFree(ctx->d_sboxes); cudaFree(ctx->d_sorder); cudaFree(ctx->d_mask);
  cudaFree(ctx->d_fbox); cudaFree(ctx->d_fcls); cudaFree(ctx->d_fscore);
  cudaFree(ctx->d_fidx); cudaFree(ctx->d_fcount); cudaFree(ctx->d_dmask);
  cudaFree(ctx->d_roi_level);
  roialign_release(ctx);
  cudaFree(ctx->d_gather_send); cudaFree(ctx->d_gather_recv);
  for (auto& pr : ctx->prof_pool) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int mrcnn_set_stream(mrcnn_ctx* ctx, void* cuda_stream) {
  if (!ctx) return MRCNN_EINVAL;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return MRCNN_OK;
}

int mrcnn_synchronize(mrcnn_ctx* ctx) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return MRCNN_OK;
}

int mrcnn_set_anchors(mrcnn_ctx* ctx, const float* anchors, int64_t n) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, anchors && n >= 1, "set_anchors: null or empty");
  MRCNN_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->d_anchors); ctx->d_anchors = nullptr;
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_anchors, sizeof(float) * 4 * n));
  MRCNN_CUDA_TRY(ctx, cudaMemcpy(ctx->d_anchors, anchors, sizeof(float) * 4 * n, cudaMemcpyDefault));
  ctx->num_anchors = n;
  return MRCNN_OK;
}

int64_t mrcnn_num_anchors(const mrcnn_ctx* ctx) { return ctx ? ctx->num_anchors : 0; }

int mrcnn_set_weights(mrcnn_ctx* ctx, int which, const void* blob, size_t bytes) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, which >= 0 && which <= 2 && blob && bytes > 0, "set_weights: bad argument");
  return dense_load_weights(ctx, which, blob, bytes);
}

int64_t mrcnn_launch_count(const mrcnn_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mrcnn_last_stage_times(const mrcnn_ctx* ctx, int max_stages, const char** names_out, float* ms_out) {
  if (!ctx) return 0;
  dense_collect_stage_times(const_cast<mrcnn_ctx*>(ctx));
  int n = 0;
  for (auto& kv : ctx->stage_ms) {
    if (n >= max_stages) break;
    if (names_out) names_out[n] = kv.first;
    if (ms_out) ms_out[n] = kv.second;
    ++n;
  }
  return n;
}

static const char* kProfNames[PROF_NUM_CLASSES] = {
    "conv_gemm_tcgen05", "roialign", "topk_radix_select", "sort_decode", "nms_iou_mask", "nms_resolve",
    "detection_filter", "detection_finalize", "glue"};

int mrcnn_profile_enable(mrcnn_ctx* ctx, int on) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->profiling = on != 0;
  ctx->prof_recs.clear();
  ctx->prof_used = 0;
  return MRCNN_OK;
}

int mrcnn_profile_read(mrcnn_ctx* ctx, int max_classes, const char** names_out, float* ms_out,
                       int64_t* launches_out, double* work_out) {
  if (!ctx) return 0;
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return 0;
  float ms[PROF_NUM_CLASSES] = {0};
  int64_t cnt[PROF_NUM_CLASSES] = {0};
  double work[PROF_NUM_CLASSES] = {0};
  for (auto& r : ctx->prof_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) { ms[r.cls] += t; cnt[r.cls] += r.launches; work[r.cls] += r.work; }
  }
  ctx->prof_recs.clear();
  ctx->prof_used = 0;
  int n = PROF_NUM_CLASSES < max_classes ? PROF_NUM_CLASSES : max_classes;
  for (int i = 0; i < n; ++i) {
    if (names_out) names_out[i] = kProfNames[i];
    if (ms_out) ms_out[i] = ms[i];
    if (launches_out) launches_out[i] = cnt[i];
    if (work_out) work_out[i] = work[i];
  }
  return n;
}

// ---- ProposalLayer -----------------------------------------------------------
int mrcnn_proposal_output_shape(const mrcnn_ctx* ctx, int64_t shape_out[2]) {
  if (!ctx || !shape_out) return MRCNN_EINVAL;
  shape_out[0] = ctx->cfg.max_proposals; shape_out[1] = 4;     // ProposalLayer.swift:97-101
  return MRCNN_OK;
}

int mrcnn_proposal_eval(mrcnn_ctx* ctx, int batch, int64_t N, const float* probs, const float* deltas,
                        float* rois_out, int32_t* keep_anchor_out, int32_t* count_out) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, probs && deltas && rois_out, "proposal_eval: null pointer");
  MRCNN_REQUIRE(ctx, batch >= 1 && N >= 1, "proposal_eval: bad batch / num_anchors");
  cudaSetDevice(ctx->device);
  Stager st(ctx);
  int rc = MRCNN_OK;
  const int mp = ctx->cfg.max_proposals;
  const float* dp = (const float*)st.in(probs, sizeof(float) * 2 * N * batch, &rc);
  const float* dd = (const float*)st.in(deltas, sizeof(float) * 4 * N * batch, &rc);
  float* dr = (float*)st.out(rois_out, sizeof(float) * 4 * mp * batch, &rc);
  int32_t* dk = (int32_t*)st.out(keep_anchor_out, sizeof(int32_t) * mp * batch, &rc);
  int32_t* dc = (int32_t*)st.out(count_out, sizeof(int32_t) * batch, &rc);
  if (rc) return mrcnn_fail(ctx, rc, "proposal_eval: staging failed");
  rc = proposal_run(ctx, batch, N, dp, dd, dr, dk, dc);
  if (rc) return rc;
  return st.finish();
}

// ---- PyramidROIAlignLayer -------------------------------------------------------
int mrcnn_pyramid_roialign_output_shape(const mrcnn_ctx* ctx, int64_t R, int64_t C, int pool, int64_t shape_out[4]) {
  if (!ctx || !shape_out) return MRCNN_EINVAL;
  shape_out[0] = R; shape_out[1] = C; shape_out[2] = pool; shape_out[3] = pool;   // PyramidROIAlignLayer.swift:65-77
  return MRCNN_OK;
}

int mrcnn_pyramid_roialign_eval(mrcnn_ctx* ctx, int batch, const float* rois, int roi_row_stride, int64_t R,
                                const float* const fmaps[4], const int32_t hw[8], int64_t C, int pool,
                                float* out, int32_t* level_out) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, rois && fmaps && hw && out, "roialign_eval: null pointer");
  MRCNN_REQUIRE(ctx, batch >= 1 && R >= 1 && C >= 1 && pool >= 1 && pool <= 64 && roi_row_stride >= 4, "roialign_eval: bad sizes");
  for (int l = 0; l < 4; ++l)
    MRCNN_REQUIRE(ctx, fmaps[l] && hw[2 * l] >= 1 && hw[2 * l + 1] >= 1, "roialign_eval: null feature map / bad map size");
  cudaSetDevice(ctx->device);
  Stager st(ctx);
  int rc = MRCNN_OK;
  const float* dr = (const float*)st.in(rois, sizeof(float) * roi_row_stride * R * batch, &rc);
  const float* dm[4];
  for (int l = 0; l < 4; ++l) {
    MRCNN_REQUIRE(ctx, fmaps[l], "roialign_eval: null feature map");
    dm[l] = (const float*)st.in(fmaps[l], sizeof(float) * C * hw[2 * l] * hw[2 * l + 1] * batch, &rc);
  }
  float* dout = (float*)st.out(out, sizeof(float) * R * C * pool * pool * batch, &rc);
  int32_t* dl = (int32_t*)st.out(level_out, sizeof(int32_t) * R * batch, &rc);
  if (rc) return mrcnn_fail(ctx, rc, "roialign_eval: staging failed");
  rc = roialign_chw_run(ctx, batch, dr, roi_row_stride, R, dm, hw, C, pool, dout, dl);
  if (rc) return rc;
  return st.finish();
}

int mrcnn_roialign_nhwc_f16(mrcnn_ctx* ctx, int batch, const float* rois, int roi_row_stride, int64_t R,
                            const void* const fmaps[4], const int32_t hw[8], int64_t C, int pool,
                            void* out, int32_t* level_out) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, rois && fmaps && hw && out, "roialign_nhwc: null pointer");
  MRCNN_REQUIRE(ctx, batch >= 1 && R >= 1 && C >= 8 && (C % 8) == 0 && pool >= 1 && pool <= 64 && roi_row_stride >= 4,
                "roialign_nhwc: bad sizes (channels must be a multiple of 8, 1 <= pool <= 64)");
  for (int l = 0; l < 4; ++l)      // checked before anything is staged: a NULL level would fault on the device
    MRCNN_REQUIRE(ctx, fmaps[l] && hw[2 * l] >= 1 && hw[2 * l + 1] >= 1, "roialign_nhwc: null feature map / bad map size");
  cudaSetDevice(ctx->device);
  Stager st(ctx);
  int rc = MRCNN_OK;
  const float* dr = (const float*)st.in(rois, sizeof(float) * roi_row_stride * R * batch, &rc);
  const __half* dm[4];
  for (int l = 0; l < 4; ++l)
    dm[l] = (const __half*)st.in(fmaps[l], sizeof(__half) * C * hw[2 * l] * hw[2 * l + 1] * batch, &rc);
  __half* dout = (__half*)st.out(out, sizeof(__half) * R * C * pool * pool * batch, &rc);
  int32_t* dl = (int32_t*)st.out(level_out, sizeof(int32_t) * R * batch, &rc);
  if (rc) return mrcnn_fail(ctx, rc, "roialign_nhwc: staging failed");
  rc = roialign_nhwc_f16_run(ctx, batch, dr, roi_row_stride, R, dm, hw, C, pool, dout, dl);
  if (rc) return rc;
  return st.finish();
}

// ---- TimeDistributedClassifierLayer (post-processing half) ------------------------
int mrcnn_classifier_select(mrcnn_ctx* ctx, int batch, int64_t R, const float* probabilities,
                            const float* bounding_boxes, float* out) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, probabilities && bounding_boxes && out, "classifier_select: null pointer");
  MRCNN_REQUIRE(ctx, batch >= 1 && R >= 1, "classifier_select: bad sizes");
  cudaSetDevice(ctx->device);
  const int ncls = ctx->cfg.num_classes;
  Stager st(ctx);
  int rc = MRCNN_OK;
  const float* dp = (const float*)st.in(probabilities, sizeof(float) * ncls * R * batch, &rc);
  const float* db = (const float*)st.in(bounding_boxes, sizeof(float) * ncls * 4 * R * batch, &rc);
  float* dout = (float*)st.out(out, sizeof(float) * 6 * R * batch, &rc);
  if (rc) return mrcnn_fail(ctx, rc, "classifier_select: staging failed");
  rc = classifier_select_run(ctx, batch, R, ncls, dp, db, dout);
  if (rc) return rc;
  return st.finish();
}

// ---- DetectionLayer ------------------------------------------------------------
int mrcnn_detection_output_shape(const mrcnn_ctx* ctx, int64_t shape_out[2]) {
  if (!ctx || !shape_out) return MRCNN_EINVAL;
  shape_out[0] = ctx->cfg.max_detections; shape_out[1] = 6;     // DetectionLayer.swift:94-105
  return MRCNN_OK;
}

int mrcnn_detection_eval(mrcnn_ctx* ctx, int batch, int64_t R, const float* rois, const float* classifications,
                         float* out, int32_t* keep_roi_out, int32_t* count_out) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, rois && classifications && out, "detection_eval: null pointer");
  MRCNN_REQUIRE(ctx, batch >= 1 && R >= 1, "detection_eval: bad sizes");
  cudaSetDevice(ctx->device);
  const int md = ctx->cfg.max_detections;
  Stager st(ctx);
  int rc = MRCNN_OK;
  const float* dr = (const float*)st.in(rois, sizeof(float) * 4 * R * batch, &rc);
  const float* dc = (const float*)st.in(classifications, sizeof(float) * 6 * R * batch, &rc);
  float* dout = (float*)st.out(out, sizeof(float) * 6 * md * batch, &rc);
  int32_t* dk = (int32_t*)st.out(keep_roi_out, sizeof(int32_t) * md * batch, &rc);
  int32_t* dn = (int32_t*)st.out(count_out, sizeof(int32_t) * batch, &rc);
  if (rc) return mrcnn_fail(ctx, rc, "detection_eval: staging failed");
  rc = detection_run(ctx, batch, R, dr, dc, dout, dk, dn);
  if (rc) return rc;
  return st.finish();
}

// ---- Detection.swift decode -------------------------------------------------------
int mrcnn_detections_decode(mrcnn_ctx* ctx, int batch, const float* detections, const float* masks,
                            int32_t* count_out, int32_t* index_out, double* bbox_out, int32_t* class_out,
                            double* score_out, uint8_t* mask_u8_out) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, detections && count_out && index_out && bbox_out && class_out && score_out,
                "detections_decode: null pointer");
  cudaSetDevice(ctx->device);
  const int D = ctx->cfg.max_detections, S = 2 * ctx->cfg.pool_size_mask;
  Stager st(ctx);
  int rc = MRCNN_OK;
  const float* dd = (const float*)st.in(detections, sizeof(float) * 6 * D * batch, &rc);
  const float* dm = (const float*)st.in(masks, sizeof(float) * (size_t)S * S * D * batch, &rc);
  int32_t* dc = (int32_t*)st.out(count_out, sizeof(int32_t) * batch, &rc);
  int32_t* di = (int32_t*)st.out(index_out, sizeof(int32_t) * D * batch, &rc);
  double* db = (double*)st.out(bbox_out, sizeof(double) * 4 * D * batch, &rc);
  int32_t* dk = (int32_t*)st.out(class_out, sizeof(int32_t) * D * batch, &rc);
  double* ds = (double*)st.out(score_out, sizeof(double) * D * batch, &rc);
  uint8_t* du = (uint8_t*)st.out(mask_u8_out, (size_t)S * S * D * batch, &rc);
  if (rc) return mrcnn_fail(ctx, rc, "detections_decode: staging failed");
  rc = detections_decode_run(ctx, batch, D, S, dd, dm, dc, di, db, dk, ds, du);
  if (rc) return rc;
  return st.finish();
}

}  // extern "C"
