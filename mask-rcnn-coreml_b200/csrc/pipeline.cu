// pipeline.cu -- the dense graph of the Mask-RCNN model split (MaskRCNN / Classifier /
// Mask, README.md:107-116 of the reference) and the fused prediction pipeline.
//
// The reference runs these graphs inside Core ML from .mlmodel artefacts that are
// not part of its source tree; the architecture here is the
// credited Matterport one: ResNet101/50 + FPN + RPN ("main"), the TimeDistributed
// classifier head and the mask head.  Every dense layer is one launch of the
// tcgen05 implicit-GEMM kernel (conv_gemm.cuh) with BN folded into weight + bias
// and bias / ReLU / residual / FPN top-down add / pixel-shuffle fused in the
// epilogue; everything else is a small memory-bound kernel from aux_kernels.cuh or
// one of the custom-layer kernels (proposal.cu, roialign.cu, detection.cu).
//
//   image u8 ─ preprocess ─ ResNet ─ FPN ─ RPN ─ ProposalLayer ─ PyramidROIAlign(7)
//     ─ classifier head ─ DetectionLayer ─ PyramidROIAlign(14) ─ mask head
//
// Activations are NHWC fp16 (fp32 accumulate); RPN / classifier outputs that feed
// the exact-arithmetic custom layers are fp32.
#include "dense.h"
#include "aux_kernels.cuh"
#include <dlfcn.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <memory>

// ------------------------------------------------------------------------------
// Weight blobs ("MRCNNW1"): header, tensor table, 256-byte aligned payload.
// Written by mask-rcnn-coreml_b200/weights.py (pack_blob); DESIGN.md "Weight blobs".
// ------------------------------------------------------------------------------
struct BlobHeader { char magic[8]; uint32_t version, which, n_tensors, reserved; };
struct BlobEntry { char name[64]; uint32_t dtype, ndim; uint64_t dims[4]; uint64_t offset, nbytes; };

struct WTensor { const void* d = nullptr; int dtype = 0, ndim = 0; int64_t dims[4] = {0, 0, 0, 0}; size_t bytes = 0; };
struct WeightSet { void* d_base = nullptr; size_t bytes = 0; std::map<std::string, WTensor> t; bool loaded = false; };

struct Buf { void* p = nullptr; size_t bytes = 0; };
typedef std::vector<std::function<int(mrcnn_ctx*)>> Graph;

struct DenseModel {
  WeightSet ws[3];
  std::map<std::string, Buf> bufs;
  std::map<int, std::shared_ptr<Graph>> g_backbone;       // keyed by 2 * batch + parity (output set, see build_backbone)
  std::map<int64_t, std::shared_ptr<Graph>> g_cls, g_mask; // keyed by total rois / detections
  std::map<int64_t, std::shared_ptr<ConvPlan>> mask_last;  // deconv + fused mask tail (output pointer patched per call)
  int lvl_h[5] = {0}, lvl_w[5] = {0};
  int64_t n_anchors = 0;
  const uint8_t* rgb = nullptr;      // input of the running predict (device pointer, not owned)
  cudaEvent_t ev[16] = {nullptr};
  int n_ev = 0;
  const char* ev_name[16] = {nullptr};
  // ---- streaming prediction (mrcnn_predict_submit / mrcnn_predict_wait): two batches in flight ----
  struct StreamSlot { cudaEvent_t h2d_done = nullptr, computed = nullptr, done = nullptr; };
  StreamSlot slot[2];
  cudaStream_t copy_stream = nullptr;   // H2D of batch i+1 runs here while batch i computes on ctx->stream
  cudaStream_t copy_out_stream = nullptr;   // D2H of batch i runs here while batch i+1 computes
  cudaStream_t heads_stream = nullptr;      // proposal .. mask section of batch i runs here under the backbone of batch i+1
  cudaEvent_t backbone_done[2] = {nullptr, nullptr};
  uint64_t submitted = 0, completed = 0;
};

static DenseModel* model_of(mrcnn_ctx* ctx) {
  if (!ctx->dense) ctx->dense = new DenseModel();
  return ctx->dense;
}

static int get_buf(mrcnn_ctx* ctx, const char* name, size_t bytes, void** out) {
  DenseModel* m = model_of(ctx);
  Buf& b = m->bufs[name];
  if (b.bytes < bytes) {
    if (b.p) {
      // growing a buffer would invalidate the tensor maps of the cached graphs
      MRCNN_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      cudaFree(b.p); b.p = nullptr;
      m->g_backbone.clear(); m->g_cls.clear(); m->g_mask.clear(); m->mask_last.clear();
    }
    MRCNN_CUDA_TRY(ctx, cudaMalloc(&b.p, bytes));
    b.bytes = bytes;
  }
  *out = b.p;
  return MRCNN_OK;
}

int dense_load_weights(mrcnn_ctx* ctx, int which, const void* blob, size_t bytes) {
  DenseModel* m = model_of(ctx);
  MRCNN_REQUIRE(ctx, bytes >= sizeof(BlobHeader), "set_weights: blob too small");
  BlobHeader h;
  memcpy(&h, blob, sizeof(h));
  if (memcmp(h.magic, "MRCNNW1\0", 8) != 0 || h.version != 1)
    return mrcnn_fail(ctx, MRCNN_EIO, "set_weights: bad magic / version (expected MRCNNW1 v1)");
  if ((int)h.which != which) return mrcnn_fail(ctx, MRCNN_EIO, "set_weights: blob is for a different model slot");
  const size_t table = sizeof(BlobHeader) + (size_t)h.n_tensors * sizeof(BlobEntry);
  if (table > bytes) return mrcnn_fail(ctx, MRCNN_EIO, "set_weights: truncated tensor table");
  MRCNN_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  WeightSet& w = m->ws[which];
  cudaFree(w.d_base); w.d_base = nullptr; w.t.clear(); w.loaded = false;
  m->g_backbone.clear(); m->g_cls.clear(); m->g_mask.clear(); m->mask_last.clear();
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&w.d_base, bytes));
  MRCNN_CUDA_TRY(ctx, cudaMemcpy(w.d_base, blob, bytes, cudaMemcpyHostToDevice));
  w.bytes = bytes;
  const BlobEntry* e = (const BlobEntry*)((const char*)blob + sizeof(BlobHeader));
  for (uint32_t i = 0; i < h.n_tensors; ++i) {
    BlobEntry en;
    memcpy(&en, &e[i], sizeof(en));
    en.name[63] = 0;
    if (en.offset % 16 != 0 || en.offset + en.nbytes > bytes || en.ndim > 4)
      return mrcnn_fail(ctx, MRCNN_EIO, std::string("set_weights: bad table entry ") + en.name);
    WTensor t;
    t.d = (const char*)w.d_base + en.offset; t.dtype = (int)en.dtype; t.ndim = (int)en.ndim; t.bytes = en.nbytes;
    for (int k = 0; k < 4; ++k) t.dims[k] = (int64_t)en.dims[k];
    w.t[en.name] = t;
  }
  w.loaded = true;
  return MRCNN_OK;
}

void dense_destroy(mrcnn_ctx* ctx) {
  DenseModel* m = ctx->dense;
  if (!m) return;
  for (auto& kv : m->bufs) cudaFree(kv.second.p);
  for (int i = 0; i < 3; ++i) cudaFree(m->ws[i].d_base);
  for (int i = 0; i < 16; ++i) if (m->ev[i]) cudaEventDestroy(m->ev[i]);
  for (int i = 0; i < 2; ++i) {
    if (m->slot[i].h2d_done) cudaEventDestroy(m->slot[i].h2d_done);
    if (m->slot[i].computed) cudaEventDestroy(m->slot[i].computed);
    if (m->slot[i].done) cudaEventDestroy(m->slot[i].done);
  }
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  if (m->copy_out_stream) cudaStreamDestroy(m->copy_out_stream);
  if (m->heads_stream) cudaStreamDestroy(m->heads_stream);
  for (int i = 0; i < 2; ++i) if (m->backbone_done[i]) cudaEventDestroy(m->backbone_done[i]);
  delete m;
  ctx->dense = nullptr;
}

// ------------------------------------------------------------------------------
// Graph building helpers
// ------------------------------------------------------------------------------
struct ConvArgs {
  const char* wname = nullptr;
  const __half* x = nullptr;
  int n = 1, h = 1, w = 1, cin = 64, ld_in = 0;
  int cout = 64, k = 1, stride = 1, pad = 0, relu = 0;
  const __half* res = nullptr;
  int res_mode = 0, res_h = 0, res_w = 0, res_ld = 0;
  void* out = nullptr;
  int out_f32 = 0, ldc = 0, deconv_c = 0;
  int bn = 0;
};

static int find_w(mrcnn_ctx* ctx, int which, const std::string& name, int dtype, WTensor* out) {
  DenseModel* m = model_of(ctx);
  auto it = m->ws[which].t.find(name);
  if (it == m->ws[which].t.end()) return mrcnn_fail(ctx, MRCNN_EIO, "weights: tensor '" + name + "' missing from the blob");
  if (it->second.dtype != dtype) return mrcnn_fail(ctx, MRCNN_EIO, "weights: tensor '" + name + "' has the wrong dtype");
  *out = it->second;
  return MRCNN_OK;
}

static int make_conv_launch(mrcnn_ctx* ctx, int which, const ConvArgs& a, ConvLaunch* out) {
  WTensor w, b;
  int rc = find_w(ctx, which, std::string(a.wname) + ".w", 0, &w);
  if (rc) return rc;
  rc = find_w(ctx, which, std::string(a.wname) + ".b", 1, &b);
  if (rc) return rc;
  if (w.ndim != 4 || w.dims[0] != a.cout || w.dims[1] != a.k || w.dims[2] != a.k || w.dims[3] != a.cin || b.dims[0] != a.cout) {
    char msg[256];
    snprintf(msg, sizeof(msg), "weights: '%s' has shape [%lld,%lld,%lld,%lld], expected [%d,%d,%d,%d]", a.wname,
             (long long)w.dims[0], (long long)w.dims[1], (long long)w.dims[2], (long long)w.dims[3], a.cout, a.k, a.k, a.cin);
    return mrcnn_fail(ctx, MRCNN_EIO, msg);
  }
  ConvLaunch L;
  L.x = a.x; L.n = a.n; L.h_in = a.h; L.w_in = a.w; L.cin = a.cin; L.ld_in = a.ld_in;
  L.w = (const __half*)w.d; L.cout = a.cout; L.kh = a.k; L.kw = a.k; L.stride = a.stride; L.pad = a.pad;
  L.bias = (const float*)b.d; L.residual = a.res; L.res_mode = a.res ? a.res_mode : 0;
  L.res_h = a.res_h; L.res_w = a.res_w; L.res_ld = a.res_ld;
  L.relu = a.relu; L.out_f32 = a.out_f32; L.out = a.out; L.ldc = a.ldc; L.bn = a.bn;
  if (a.deconv_c) { L.deconv = 1; L.deconv_c = a.deconv_c; }
  *out = L;
  return MRCNN_OK;
}

static int add_conv_launch(mrcnn_ctx* ctx, Graph& g, const ConvLaunch& L, const char* name) {
  auto plan = std::make_shared<ConvPlan>();
  int rc = conv_plan_build(ctx, L, plan.get());
  if (rc) { ctx->err = std::string(name) + ": " + ctx->err; return rc; }
  g.push_back([plan](mrcnn_ctx* c) { return conv_plan_run(c, *plan); });
  return MRCNN_OK;
}

static int add_conv(mrcnn_ctx* ctx, Graph& g, int which, const ConvArgs& a) {
  ConvLaunch L;
  int rc = make_conv_launch(ctx, which, a, &L);
  if (rc) return rc;
  return add_conv_launch(ctx, g, L, a.wname);
}

// A ResNet stage: its layers, launched one by one.  (Two experiments that regrouped these launches lost on B200 and were
// removed: one persistent dataflow launch per stage, and per-sub-batch launches to keep a block in L2 -- DESIGN.md 8.)
// A block's 1x1 expansion (+ residual) followed by the next block's 1x1 reduction is ONE launch where the fused kernel
// supports the shapes (csrc/conv_fused.cuh; MRCNN_CONV_FUSE=0 launches them separately, results are bit-identical).
static int add_stage(mrcnn_ctx* ctx, Graph& g, const std::vector<ConvLaunch>& layers) {
  static int env_fuse = -1;
  if (env_fuse < 0) { const char* e = getenv("MRCNN_CONV_FUSE"); env_fuse = e ? atoi(e) : 1; }
  for (size_t i = 0; i < layers.size(); ++i) {
    if (env_fuse && i + 1 < layers.size() && fused_plan_supported(layers[i], layers[i + 1])) {
      auto plan = std::make_shared<FusedPlan>();
      int rc = fused_plan_build(ctx, layers[i], layers[i + 1], plan.get());
      if (rc) { ctx->err = "resnet stage, fused expansion + reduction: " + ctx->err; return rc; }
      g.push_back([plan](mrcnn_ctx* c) { return fused_plan_run(c, *plan); });
      ++i;
      continue;
    }
    int rc = add_conv_launch(ctx, g, layers[i], "resnet stage layer");
    if (rc) return rc;
  }
  return MRCNN_OK;
}

#define TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

static inline int grid1d(int64_t total, int threads) { return (int)((total + threads - 1) / threads); }

// ------------------------------------------------------------------------------
// main model: ResNet + FPN + RPN
// ------------------------------------------------------------------------------
static int build_backbone(mrcnn_ctx* ctx, int B, int parity, std::shared_ptr<Graph>* out_graph) {
  DenseModel* m = model_of(ctx);
  MRCNN_REQUIRE(ctx, m->ws[0].loaded, "main model weights not loaded (mrcnn_set_weights(ctx, 0, ...) / main_model_path)");
  const mrcnn_config& cfg = ctx->cfg;
  const int H = cfg.image_h, W = cfg.image_w;
  MRCNN_REQUIRE(ctx, H % 64 == 0 && W % 64 == 0 && H >= 128 && W >= 128, "image size must be a multiple of 64 (>= 128)");
  MRCNN_REQUIRE(ctx, cfg.architecture == 101 || cfg.architecture == 50, "architecture must be 101 or 50");
  const int MB = cfg.max_batch > B ? cfg.max_batch : B;     // buffers are sized for max_batch
  auto g = std::make_shared<Graph>();
  const int Hs = (H + 6) / 2, Ws = (W + 6) / 2;             // zero-padded (3) + space-to-depth(2)
  const int H1 = H / 2, W1 = W / 2;                         // conv1 output
  const int H2 = H / 4, W2 = W / 4;                         // C2
  auto elems = [&](int h, int w, int c) { return (size_t)MB * h * w * c; };
  __half *s2d, *a, *b, *pool, *t1, *t2, *sc, *c2, *c3, *c4, *c5;
  TRY(get_buf(ctx, "s2d", elems(Hs, Ws, 16) * 2, (void**)&s2d));
  TRY(get_buf(ctx, "actA", elems(H1, W1, 64) * 2, (void**)&a));
  TRY(get_buf(ctx, "actB", elems(H2, W2, 256) * 2, (void**)&b));
  TRY(get_buf(ctx, "pool", elems(H2, W2, 64) * 2, (void**)&pool));
  TRY(get_buf(ctx, "t1", elems(H2, W2, 64) * 2, (void**)&t1));
  TRY(get_buf(ctx, "t2", elems(H2, W2, 64) * 2, (void**)&t2));
  TRY(get_buf(ctx, "sc", elems(H2, W2, 256) * 2, (void**)&sc));
  TRY(get_buf(ctx, "c2", elems(H2, W2, 256) * 2, (void**)&c2));
  TRY(get_buf(ctx, "c3", elems(H2 / 2, W2 / 2, 512) * 2, (void**)&c3));
  TRY(get_buf(ctx, "c4", elems(H2 / 4, W2 / 4, 1024) * 2, (void**)&c4));
  TRY(get_buf(ctx, "c5", elems(H2 / 8, W2 / 8, 2048) * 2, (void**)&c5));

  // ---- stem: preprocess (mean subtraction, Conversion/task.py:73-75) + conv1 7x7/2 as a 4-tap GEMM
  {
    const float m0 = cfg.mean_rgb[0], m1 = cfg.mean_rgb[1], m2 = cfg.mean_rgb[2];
    g->push_back([=](mrcnn_ctx* c) -> int {
      const uint8_t* rgb = c->dense->rgb;   // set by the caller (device pointer)
      ProfScope ps(c, PROF_GLUE, (double)B * (H * W * 3.0 + Hs * Ws * 32.0));
      preprocess_s2d_kernel<<<grid1d((int64_t)B * Hs * Ws, 256), 256, 0, c->stream>>>(rgb, B, H, W, Hs, Ws, m0, m1, m2, s2d);
      MRCNN_LAUNCH_CHECK(c);
      return MRCNN_OK;
    });
    WTensor w, bb;
    TRY(find_w(ctx, 0, "conv1.w", 0, &w));
    TRY(find_w(ctx, 0, "conv1.b", 1, &bb));
    MRCNN_REQUIRE(ctx, w.dims[0] == 64 && w.dims[1] == 4 && w.dims[2] == 1 && w.dims[3] == 64, "conv1.w must be [64,4,1,64] (space-to-depth packed)");
    ConvLaunch L;
    L.x = s2d; L.n = B; L.h_in = Hs; L.w_in = Ws - 3; L.cin = 64;
    L.custom_view = true;
    L.a_dims[0] = 64; L.a_dims[1] = (uint64_t)(Ws - 3); L.a_dims[2] = (uint64_t)Hs; L.a_dims[3] = (uint64_t)B;
    L.a_strides[0] = 32; L.a_strides[1] = (uint64_t)Ws * 32; L.a_strides[2] = (uint64_t)Hs * Ws * 32;
    L.w = (const __half*)w.d; L.cout = 64; L.kh = 4; L.kw = 1; L.stride = 1; L.pad = 0;
    L.ntaps_override = 4;
    for (int t = 0; t < 4; ++t) { L.tap_dx[t] = 0; L.tap_dy[t] = (int8_t)t; }
    L.bias = (const float*)bb.d; L.relu = 1; L.out = a; L.h_out = H1; L.w_out = W1;
    auto plan = std::make_shared<ConvPlan>();
    TRY(conv_plan_build(ctx, L, plan.get()));
    plan->flops = 2.0 * B * H1 * W1 * 64.0 * 147.0;        // useful flops (the zero-padded taps do not count)
    g->push_back([plan](mrcnn_ctx* c) { return conv_plan_run(c, *plan); });
  }
  // ---- max pool 3x3/2 -> [B, H2, W2, 64]
  g->push_back([=](mrcnn_ctx* c) -> int {
    ProfScope ps(c, PROF_GLUE, (double)B * (H1 * W1 + H2 * W2) * 128.0);
    maxpool3x3s2_kernel<<<grid1d((int64_t)B * H2 * W2 * 8, 256), 256, 0, c->stream>>>(a, B, H1, W1, 64, H2, W2, pool);
    MRCNN_LAUNCH_CHECK(c);
    return MRCNN_OK;
  });
  // ---- residual stages
  const int nblocks101[4] = {3, 4, 23, 3}, nblocks50[4] = {3, 4, 6, 3};
  const int* nb = cfg.architecture == 101 ? nblocks101 : nblocks50;
  __half* x = pool;         // current block input
  __half* pp[2] = {a, b};   // ping-pong block outputs (a is free again after the max pool)
  int cur = 1;              // x lives in pp[cur]
  int h = H2, w = W2, cin = 64;
  __half* stage_out[4] = {c2, c3, c4, c5};
  for (int s = 0; s < 4; ++s) {
    const int f = 64 << s;
    std::vector<ConvLaunch> stage_layers;
    auto stage_conv = [&](const ConvArgs& ca) -> int {
      ConvLaunch L;
      int rc = make_conv_launch(ctx, 0, ca, &L);
      if (rc) return rc;
      stage_layers.push_back(L);
      return MRCNN_OK;
    };
    for (int i = 0; i < nb[s]; ++i) {
      const int stride = (i == 0 && s > 0) ? 2 : 1;
      const int ho = h / stride, wo = w / stride;
      char n2a[64], n2b[64], n2c[64], n1[64];
      snprintf(n2a, 64, "res%d.%d.2a", s + 2, i); snprintf(n2b, 64, "res%d.%d.2b", s + 2, i);
      snprintf(n2c, 64, "res%d.%d.2c", s + 2, i); snprintf(n1, 64, "res%d.%d.1", s + 2, i);
      __half* y = (i == nb[s] - 1) ? stage_out[s] : pp[cur ^ 1];
      ConvArgs A;
      A.wname = n2a; A.x = x; A.n = B; A.h = h; A.w = w; A.cin = cin; A.cout = f; A.k = 1; A.stride = stride; A.relu = 1; A.out = t1;
      TRY(stage_conv(A));
      ConvArgs Bc;
      Bc.wname = n2b; Bc.x = t1; Bc.n = B; Bc.h = ho; Bc.w = wo; Bc.cin = f; Bc.cout = f; Bc.k = 3; Bc.pad = 1; Bc.relu = 1; Bc.out = t2;
      TRY(stage_conv(Bc));
      const __half* res = x;
      if (i == 0) {
        ConvArgs S;
        S.wname = n1; S.x = x; S.n = B; S.h = h; S.w = w; S.cin = cin; S.cout = 4 * f; S.k = 1; S.stride = stride; S.out = sc;
        TRY(stage_conv(S));
        res = sc;
      }
      ConvArgs Cc;
      Cc.wname = n2c; Cc.x = t2; Cc.n = B; Cc.h = ho; Cc.w = wo; Cc.cin = f; Cc.cout = 4 * f; Cc.k = 1; Cc.relu = 1;
      Cc.res = res; Cc.res_mode = 1; Cc.out = y;
      TRY(stage_conv(Cc));
      if (y == pp[cur ^ 1]) cur ^= 1;
      x = y; h = ho; w = wo; cin = 4 * f;
    }
    TRY(add_stage(ctx, *g, stage_layers));
  }
  // ---- FPN
  const int lh[5] = {H2, H2 / 2, H2 / 4, H2 / 8, H2 / 16}, lw[5] = {W2, W2 / 2, W2 / 4, W2 / 8, W2 / 16};
  for (int l = 0; l < 5; ++l) { m->lvl_h[l] = lh[l]; m->lvl_w[l] = lw[l]; }
  __half *mrg[4], *p[5];
  const char* mn[4] = {"m2", "m3", "m4", "m5"};
  // what the stage AFTER the backbone reads (P2..P5, RPN outputs) exists twice: in the streaming API the backbone of
  // batch i + 1 runs while the proposal / heads section of batch i still reads its set (parity = batch & 1)
  const char* pn0[5] = {"p2", "p3", "p4", "p5", "p6"};
  const char* pn1[5] = {"p2#1", "p3#1", "p4#1", "p5#1", "p6"};
  const char* const* pn = parity ? pn1 : pn0;
  for (int l = 0; l < 4; ++l) TRY(get_buf(ctx, mn[l], elems(lh[l], lw[l], 256) * 2, (void**)&mrg[l]));
  for (int l = 0; l < 5; ++l) TRY(get_buf(ctx, pn[l], elems(lh[l], lw[l], 256) * 2, (void**)&p[l]));
  const char* lat[4] = {"fpn.c2p2", "fpn.c3p3", "fpn.c4p4", "fpn.c5p5"};
  const char* outc[4] = {"fpn.p2", "fpn.p3", "fpn.p4", "fpn.p5"};
  const int cc[4] = {256, 512, 1024, 2048};
  for (int l = 3; l >= 0; --l) {
    ConvArgs A;
    A.wname = lat[l]; A.x = stage_out[l]; A.n = B; A.h = lh[l]; A.w = lw[l]; A.cin = cc[l]; A.cout = 256; A.k = 1; A.out = mrg[l];
    if (l < 3) { A.res = mrg[l + 1]; A.res_mode = 2; A.res_h = lh[l + 1]; A.res_w = lw[l + 1]; A.res_ld = 256; }   // top-down 2x nearest + add
    TRY(add_conv(ctx, *g, 0, A));
  }
  for (int l = 0; l < 4; ++l) {
    ConvArgs A;
    A.wname = outc[l]; A.x = mrg[l]; A.n = B; A.h = lh[l]; A.w = lw[l]; A.cin = 256; A.cout = 256; A.k = 3; A.pad = 1; A.out = p[l];
    TRY(add_conv(ctx, *g, 0, A));
  }
  {
    __half* p5 = p[3]; __half* p6 = p[4];
    const int h5 = lh[3], w5 = lw[3], h6 = lh[4], w6 = lw[4];
    g->push_back([=](mrcnn_ctx* c) -> int {
      ProfScope ps(c, PROF_GLUE, (double)B * h6 * w6 * 1024.0);
      subsample2_kernel<<<grid1d((int64_t)B * h6 * w6 * 32, 256), 256, 0, c->stream>>>(p5, B, h5, w5, 256, h6, w6, p6);
      MRCNN_LAUNCH_CHECK(c);
      return MRCNN_OK;
    });
  }
  // ---- RPN over P2..P6 (level-major anchor order, the order of anchors.bin)
  int64_t N = 0;
  for (int l = 0; l < 5; ++l) N += 3ll * lh[l] * lw[l];
  m->n_anchors = N;
  __half* shared; float *head, *probs, *deltas;
  TRY(get_buf(ctx, "rpn_shared", elems(lh[0], lw[0], 512) * 2, (void**)&shared));
  TRY(get_buf(ctx, "rpn_head", elems(lh[0], lw[0], 24) * 4, (void**)&head));
  TRY(get_buf(ctx, parity ? "rpn_probs#1" : "rpn_probs", (size_t)MB * N * 2 * 4, (void**)&probs));
  TRY(get_buf(ctx, parity ? "rpn_deltas#1" : "rpn_deltas", (size_t)MB * N * 4 * 4, (void**)&deltas));
  int64_t off = 0;
  for (int l = 0; l < 5; ++l) {
    ConvArgs A;
    A.wname = "rpn.shared"; A.x = p[l]; A.n = B; A.h = lh[l]; A.w = lw[l]; A.cin = 256; A.cout = 512; A.k = 3; A.pad = 1; A.relu = 1; A.out = shared;
    TRY(add_conv(ctx, *g, 0, A));
    ConvArgs Hc;
    Hc.wname = "rpn.head"; Hc.x = shared; Hc.n = B; Hc.h = lh[l]; Hc.w = lw[l]; Hc.cin = 512; Hc.cout = 18; Hc.k = 1; Hc.out = head; Hc.out_f32 = 1; Hc.ldc = 24;
    TRY(add_conv(ctx, *g, 0, Hc));
    const int hh = lh[l], ww = lw[l];
    const int64_t o = off;
    g->push_back([=](mrcnn_ctx* c) -> int {
      ProfScope ps(c, PROF_GLUE, (double)B * hh * ww * (96.0 + 72.0));
      rpn_post_kernel<<<grid1d((int64_t)B * hh * ww * 3, 256), 256, 0, c->stream>>>(head, B, hh, ww, N, o, probs, deltas);
      MRCNN_LAUNCH_CHECK(c);
      return MRCNN_OK;
    });
    off += 3ll * hh * ww;
  }
  *out_graph = g;
  return MRCNN_OK;
}

static int run_graph(mrcnn_ctx* ctx, const Graph& g) {
  for (auto& f : g) TRY(f(ctx));
  return MRCNN_OK;
}

static int backbone_graph(mrcnn_ctx* ctx, int B, std::shared_ptr<Graph>* g, int parity = 0) {
  DenseModel* m = model_of(ctx);
  auto it = m->g_backbone.find(2 * B + parity);
  if (it == m->g_backbone.end()) {
    std::shared_ptr<Graph> ng;
    TRY(build_backbone(ctx, B, parity, &ng));
    // building may have (re)allocated buffers and dropped the cache: insert afterwards
    m->g_backbone[2 * B + parity] = ng;
    *g = ng;
  } else *g = it->second;
  return MRCNN_OK;
}

// ------------------------------------------------------------------------------
// classifier head (Classifier model + TimeDistributedClassifierLayer.swift:50-88)
//   pooled [M, P, P, 256] f16 -> cls6 [M, 6] f32 (and optionally probabilities / boxes)
// ------------------------------------------------------------------------------
static int cls_graph(mrcnn_ctx* ctx, int64_t M, std::shared_ptr<Graph>* out_graph) {
  DenseModel* m = model_of(ctx);
  auto it = m->g_cls.find(M);
  if (it != m->g_cls.end()) { *out_graph = it->second; return MRCNN_OK; }
  MRCNN_REQUIRE(ctx, m->ws[1].loaded, "classifier weights not loaded (mrcnn_set_weights(ctx, 1, ...) / classifier_model_path)");
  const mrcnn_config& cfg = ctx->cfg;
  const int P = cfg.pool_size_classifier, ncls = cfg.num_classes;
  const int64_t cap = (int64_t)(cfg.max_batch > 1 ? cfg.max_batch : 1) * cfg.max_proposals;
  const int64_t MM = M > cap ? M : cap;
  const int nout = ncls * 5, ld = (nout + 7) / 8 * 8;
  __half *pooled, *f1, *f2; float* logits;
  TRY(get_buf(ctx, "pooled_cls", (size_t)MM * P * P * 256 * 2, (void**)&pooled));
  TRY(get_buf(ctx, "cls_f1", (size_t)MM * 1024 * 2, (void**)&f1));
  TRY(get_buf(ctx, "cls_f2", (size_t)MM * 1024 * 2, (void**)&f2));
  TRY(get_buf(ctx, "cls_logits", (size_t)MM * ld * 4, (void**)&logits));
  auto g = std::make_shared<Graph>();
  // conv PxP "valid" over a PxP map == one GEMM with K = P*P*256 (pooled rows are already K-major)
  WTensor w, b;
  TRY(find_w(ctx, 1, "cls.conv1.w", 0, &w));
  TRY(find_w(ctx, 1, "cls.conv1.b", 1, &b));
  MRCNN_REQUIRE(ctx, w.dims[0] == 1024 && w.dims[1] == P && w.dims[2] == P && w.dims[3] == 256, "cls.conv1.w must be [1024,P,P,256]");
  {
    ConvLaunch L;
    L.x = pooled; L.n = 1; L.h_in = 1; L.w_in = (int)M; L.cin = P * P * 256;
    L.w = (const __half*)w.d; L.cout = 1024; L.kh = 1; L.kw = 1; L.bias = (const float*)b.d; L.relu = 1; L.out = f1;
    auto plan = std::make_shared<ConvPlan>();
    TRY(conv_plan_build(ctx, L, plan.get()));
    g->push_back([plan](mrcnn_ctx* c) { return conv_plan_run(c, *plan); });
  }
  ConvArgs A;
  A.wname = "cls.conv2"; A.x = f1; A.n = 1; A.h = 1; A.w = (int)M; A.cin = 1024; A.cout = 1024; A.k = 1; A.relu = 1; A.out = f2;
  TRY(add_conv(ctx, *g, 1, A));
  ConvArgs F;
  F.wname = "cls.fc"; F.x = f2; F.n = 1; F.h = 1; F.w = (int)M; F.cin = 1024; F.cout = nout; F.k = 1; F.out = logits; F.out_f32 = 1; F.ldc = ld;
  TRY(add_conv(ctx, *g, 1, F));
  m->g_cls[M] = g;
  *out_graph = g;
  return MRCNN_OK;
}

static int cls_post(mrcnn_ctx* ctx, int64_t M, float* d_out6, float* d_probs, float* d_bbox) {
  DenseModel* m = model_of(ctx);
  const int ncls = ctx->cfg.num_classes, ld = (ncls * 5 + 7) / 8 * 8;
  const float* logits = (const float*)m->bufs["cls_logits"].p;
  ProfScope ps(ctx, PROF_GLUE, (double)M * (ld * 4.0 + 24.0));
  cls_post_kernel<<<grid1d(M * 32, 256), 256, 0, ctx->stream>>>(logits, M, ld, ncls, d_out6, d_probs, d_bbox);
  MRCNN_LAUNCH_CHECK(ctx);
  return MRCNN_OK;
}

// ------------------------------------------------------------------------------
// mask head (Mask model + TimeDistributedMaskLayer.swift:58-89)
//   pooled [M, P, P, 256] f16, valid [M], detections [M/D, D, 6] -> masks [M, 2P, 2P] f32
// ------------------------------------------------------------------------------
static int mask_graph(mrcnn_ctx* ctx, int64_t M, std::shared_ptr<Graph>* out_graph) {
  DenseModel* m = model_of(ctx);
  auto it = m->g_mask.find(M);
  if (it != m->g_mask.end()) { *out_graph = it->second; return MRCNN_OK; }
  MRCNN_REQUIRE(ctx, m->ws[2].loaded, "mask weights not loaded (mrcnn_set_weights(ctx, 2, ...) / mask_model_path)");
  const mrcnn_config& cfg = ctx->cfg;
  const int P = cfg.pool_size_mask;
  const int64_t cap = (int64_t)(cfg.max_batch > 1 ? cfg.max_batch : 1) * cfg.max_detections;
  const int64_t MM = M > cap ? M : cap;
  const bool precise = cfg.precise_masks != 0;
  const int act_c = precise ? 512 : 256;          // precise mode: activations as (hi, lo) fp16 pairs, 2*256 channels
  __half *pooled, *ma, *mb;
  int32_t *valid, *slot_valid, *slot_cls;
  TRY(get_buf(ctx, "pooled_mask", (size_t)MM * P * P * 256 * 2, (void**)&pooled));
  TRY(get_buf(ctx, "mask_a", (size_t)MM * P * P * act_c * 2, (void**)&ma));
  TRY(get_buf(ctx, "mask_b", (size_t)MM * P * P * act_c * 2, (void**)&mb));
  TRY(get_buf(ctx, "mask_valid", (size_t)MM * 4, (void**)&valid));
  TRY(get_buf(ctx, "mask_slot_valid", (size_t)MM * 4, (void**)&slot_valid));
  TRY(get_buf(ctx, "mask_slot_cls", (size_t)MM * 4, (void**)&slot_cls));
  auto g = std::make_shared<Graph>();
  const char* names[4] = {"mask.conv1", "mask.conv2", "mask.conv3", "mask.conv4"};
  // precise mode: weights of the layers that consume (hi, lo) activations, duplicated along the input channels
  // ([cout][kh][kw][cin] -> [cout][kh][kw][2*cin] = (w | w)), so that hi*w + lo*w accumulates in fp32
  auto duplicated = [&](const char* bufname, const WTensor& w, int rows, const __half** out) -> int {
    __half* d;
    TRY(get_buf(ctx, bufname, (size_t)rows * 512 * 2, (void**)&d));
    MRCNN_CUDA_TRY(ctx, cudaMemcpy2DAsync(d, 512 * 2, w.d, 256 * 2, 256 * 2, rows, cudaMemcpyDeviceToDevice, ctx->stream));
    MRCNN_CUDA_TRY(ctx, cudaMemcpy2DAsync(d + 256, 512 * 2, w.d, 256 * 2, 256 * 2, rows, cudaMemcpyDeviceToDevice, ctx->stream));
    *out = d;
    return MRCNN_OK;
  };
  const __half* x = pooled;
  int xc = 256;
  for (int i = 0; i < 4; ++i) {
    __half* y = (i & 1) ? mb : ma;
    if (!precise) {
      ConvArgs A;
      A.wname = names[i]; A.x = x; A.n = (int)M; A.h = P; A.w = P; A.cin = 256; A.cout = 256; A.k = 3; A.pad = 1; A.relu = 1; A.out = y;
      TRY(add_conv(ctx, *g, 2, A));
    } else {
      WTensor w, b;
      TRY(find_w(ctx, 2, std::string(names[i]) + ".w", 0, &w));
      TRY(find_w(ctx, 2, std::string(names[i]) + ".b", 1, &b));
      MRCNN_REQUIRE(ctx, w.dims[0] == 256 && w.dims[1] == 3 && w.dims[2] == 3 && w.dims[3] == 256, "mask.conv*.w must be [256,3,3,256]");
      const __half* wd = (const __half*)w.d;
      if (xc == 512) { char bn_[32]; snprintf(bn_, 32, "mask_wdup%d", i); TRY(duplicated(bn_, w, 256 * 9, &wd)); }
      ConvLaunch L;
      L.x = x; L.n = (int)M; L.h_in = P; L.w_in = P; L.cin = xc; L.w = wd; L.cout = 256; L.kh = 3; L.kw = 3; L.pad = 1;
      L.bias = (const float*)b.d; L.relu = 1; L.out = y; L.split_out = 1;
      auto plan = std::make_shared<ConvPlan>();
      TRY(conv_plan_build(ctx, L, plan.get()));
      g->push_back([plan](mrcnn_ctx* c) { return conv_plan_run(c, *plan); });
    }
    x = y;
    xc = act_c;
  }
  // 2x2 stride-2 transposed conv (GEMM with 4*256 outputs, one N tile per sub-pixel) + ReLU, with the final
  // class-selected 1x1 conv + sigmoid fused into its epilogue: the (M, 2P, 2P, 256) tensor is never materialised.
  WTensor w, b, wf, bf;
  TRY(find_w(ctx, 2, "mask.deconv.w", 0, &w));
  TRY(find_w(ctx, 2, "mask.deconv.b", 1, &b));
  TRY(find_w(ctx, 2, "mask.final.w", 0, &wf));
  TRY(find_w(ctx, 2, "mask.final.b", 1, &bf));
  MRCNN_REQUIRE(ctx, w.dims[0] == 1024 && w.dims[1] == 1 && w.dims[2] == 1 && w.dims[3] == 256 && b.dims[0] == 1024,
                "mask.deconv.w must be [4*256,1,1,256] with a [1024] bias (bias repeated per sub-pixel)");
  MRCNN_REQUIRE(ctx, wf.dims[0] == cfg.num_classes && wf.dims[3] == 256 && bf.dims[0] == cfg.num_classes,
                "mask.final.w must be [num_classes,1,1,256]");
  {
    const __half* wd = (const __half*)w.d;
    if (precise) TRY(duplicated("mask_wdup_deconv", w, 1024, &wd));
    ConvLaunch L;
    L.x = x; L.n = (int)M; L.h_in = P; L.w_in = P; L.cin = act_c;
    L.w = wd; L.cout = 1024; L.kh = 1; L.kw = 1; L.bias = (const float*)b.d; L.relu = 1;
    L.deconv = 1; L.deconv_c = 256; L.out = slot_valid /* patched per call */; L.out_f32 = 1; L.ldc = 256; L.bn = 256;
    L.maskdot = 1; L.md_valid = slot_valid; L.md_cls = slot_cls; L.md_w = (const __half*)wf.d; L.md_b = (const float*)bf.d;
    L.md_ncls = cfg.num_classes; L.md_precise = precise ? 1 : 0;
    auto plan = std::make_shared<ConvPlan>();
    TRY(conv_plan_build(ctx, L, plan.get()));
    plan->flops += 2.0 * M * 4 * P * P * 256.0;      // the fused class-selected 1x1
    m->mask_last[M] = plan;
  }
  m->g_mask[M] = g;
  *out_graph = g;
  return MRCNN_OK;
}

static int mask_tail(mrcnn_ctx* ctx, int batch, int D, const float* d_det, float* d_out) {
  DenseModel* m = model_of(ctx);
  const int64_t M = (int64_t)batch * D;
  auto it = m->mask_last.find(M);
  MRCNN_REQUIRE(ctx, it != m->mask_last.end(), "mask head not built for this batch");
  const int32_t* valid = (const int32_t*)m->bufs["mask_valid"].p;
  int32_t* slot_valid = (int32_t*)m->bufs["mask_slot_valid"].p;
  int32_t* slot_cls = (int32_t*)m->bufs["mask_slot_cls"].p;
  {
    ProfScope ps(ctx, PROF_GLUE, (double)M * 40.0);
    mask_slots_kernel<<<batch, 256, 0, ctx->stream>>>(valid, d_det, D, slot_valid, slot_cls);
    MRCNN_LAUNCH_CHECK(ctx);
  }
  it->second->p.out = d_out;
  return conv_plan_run(ctx, *it->second);
}

// ------------------------------------------------------------------------------
// stage timing (replaces the os_signpost intervals of the reference)
// ------------------------------------------------------------------------------
static void stage_mark(mrcnn_ctx* ctx, const char* name) {
  DenseModel* m = model_of(ctx);
  if (m->n_ev >= 16) return;
  if (!m->ev[m->n_ev]) cudaEventCreate(&m->ev[m->n_ev]);
  cudaEventRecord(m->ev[m->n_ev], ctx->stream);
  m->ev_name[m->n_ev] = name;
  m->n_ev++;
}

static void stage_collect(mrcnn_ctx* ctx) {
  DenseModel* m = model_of(ctx);
  ctx->stage_ms.clear();
  if (m->n_ev < 2) return;
  if (cudaEventSynchronize(m->ev[m->n_ev - 1]) != cudaSuccess) return;
  for (int i = 1; i < m->n_ev; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, m->ev[i - 1], m->ev[i]);
    ctx->stage_ms.push_back({m->ev_name[i], ms});
  }
}

// ------------------------------------------------------------------------------
// The fused pipeline.  All pointers are device pointers.
// ------------------------------------------------------------------------------
// Restores the context's stream when a function that switched it returns (also on the error paths).
struct StreamGuard {
  mrcnn_ctx* ctx; cudaStream_t saved;
  explicit StreamGuard(mrcnn_ctx* c) : ctx(c), saved(c->stream) {}
  ~StreamGuard() { ctx->stream = saved; }
};

// parity: which set of backbone outputs (P2..P5, RPN outputs) this batch uses.  heads: stream for everything after
// the backbone (nullptr: the context's stream) -- the streaming API runs that section of batch i beside the backbone
// of batch i + 1; it holds the latency-bound kernels (top-k, sort, NMS resolves: a few dozen CTAs) whose idle SMs the
// other batch's convolutions fill.
static int predict_device(mrcnn_ctx* ctx, int B, const uint8_t* d_rgb, float* d_det, float* d_masks,
                          int parity = 0, cudaStream_t heads = nullptr) {
  DenseModel* m = model_of(ctx);
  const mrcnn_config& cfg = ctx->cfg;
  MRCNN_REQUIRE(ctx, B >= 1, "predict: batch must be >= 1");
  MRCNN_REQUIRE(ctx, m->ws[0].loaded && m->ws[1].loaded && m->ws[2].loaded, "predict: main / classifier / mask weights must all be loaded");
  MRCNN_REQUIRE(ctx, cfg.max_proposals <= 65535, "predict: max_proposals too large");
  const int R = cfg.max_proposals, D = cfg.max_detections, P7 = cfg.pool_size_classifier, P14 = cfg.pool_size_mask;
  // build (or fetch) every graph first: building may reallocate shared buffers
  std::shared_ptr<Graph> gb, gc, gm;
  float *rois = nullptr, *cls6 = nullptr;
  const int MB = cfg.max_batch > B ? cfg.max_batch : B;
  for (int pass = 0; pass < 2; ++pass) {
    // every buffer of the call is reserved inside this loop: a buffer that grows drops the cached graphs (get_buf),
    // so the check at the end of the pass must come after the last reservation
    TRY(get_buf(ctx, "rois", (size_t)MB * R * 4 * 4, (void**)&rois));
    TRY(get_buf(ctx, "cls6", (size_t)MB * R * 6 * 4, (void**)&cls6));
    TRY(backbone_graph(ctx, B, &gb, parity));
    TRY(cls_graph(ctx, (int64_t)B * R, &gc));
    TRY(mask_graph(ctx, (int64_t)B * D, &gm));
    if (m->g_backbone.count(2 * B + parity) && m->g_cls.count((int64_t)B * R) && m->g_mask.count((int64_t)B * D) &&
        m->mask_last.count((int64_t)B * D)) break;
  }
  MRCNN_REQUIRE(ctx, m->g_backbone.count(2 * B + parity) && m->g_cls.count((int64_t)B * R) && m->g_mask.count((int64_t)B * D) &&
                m->mask_last.count((int64_t)B * D), "predict: internal error, graphs invalidated while building");
  MRCNN_REQUIRE(ctx, ctx->d_anchors && ctx->num_anchors == m->n_anchors, "predict: anchors not loaded or anchor count does not match the image size");
  m->rgb = d_rgb;     // consumed by the pre-processing closure
  m->n_ev = 0;
  stage_mark(ctx, "start");
  TRY(run_graph(ctx, *gb));
  stage_mark(ctx, "Backbone+FPN+RPN");
  StreamGuard guard(ctx);
  if (heads) {
    if (!m->backbone_done[parity]) MRCNN_CUDA_TRY(ctx, cudaEventCreateWithFlags(&m->backbone_done[parity], cudaEventDisableTiming));
    MRCNN_CUDA_TRY(ctx, cudaEventRecord(m->backbone_done[parity], ctx->stream));
    MRCNN_CUDA_TRY(ctx, cudaStreamWaitEvent(heads, m->backbone_done[parity], 0));
    ctx->stream = heads;                 // every launch below goes to the heads stream (restored by the guard)
  }
  const float* probs = (const float*)m->bufs[parity ? "rpn_probs#1" : "rpn_probs"].p;
  const float* deltas = (const float*)m->bufs[parity ? "rpn_deltas#1" : "rpn_deltas"].p;
  TRY(proposal_run(ctx, B, m->n_anchors, probs, deltas, rois, nullptr, nullptr));
  stage_mark(ctx, "Proposal-Eval");
  const char* pnames[2][4] = {{"p2", "p3", "p4", "p5"}, {"p2#1", "p3#1", "p4#1", "p5#1"}};
  const __half* fm[4];
  for (int l = 0; l < 4; ++l) fm[l] = (const __half*)m->bufs[pnames[parity][l]].p;
  int32_t hw[8];
  for (int l = 0; l < 4; ++l) { hw[2 * l] = m->lvl_h[l]; hw[2 * l + 1] = m->lvl_w[l]; }
  TRY(roialign_nhwc_f16_run(ctx, B, rois, 4, R, fm, hw, 256, P7, (__half*)m->bufs["pooled_cls"].p, nullptr));
  stage_mark(ctx, "PyramidROIAlign-Eval(7)");
  TRY(run_graph(ctx, *gc));
  TRY(cls_post(ctx, (int64_t)B * R, cls6, nullptr, nullptr));
  stage_mark(ctx, "TimeDistributedClassifierLayer-Eval");
  TRY(detection_run(ctx, B, R, rois, cls6, d_det, nullptr, nullptr));
  stage_mark(ctx, "Detection-Eval");
  int32_t* valid = (int32_t*)m->bufs["mask_valid"].p;
  TRY(roialign_nhwc_f16_run(ctx, B, d_det, 6, D, fm, hw, 256, P14, (__half*)m->bufs["pooled_mask"].p, nullptr));
  {
    ProfScope ps(ctx, PROF_GLUE, (double)B * D * 8.0);
    level_to_valid_kernel<<<grid1d((int64_t)B * D, 256), 256, 0, ctx->stream>>>(ctx->d_roi_level, (int64_t)B * D, valid);
    MRCNN_LAUNCH_CHECK(ctx);
  }
  stage_mark(ctx, "PyramidROIAlign-Eval(14)");
  TRY(run_graph(ctx, *gm));
  TRY(mask_tail(ctx, B, D, d_det, d_masks));
  stage_mark(ctx, "TimeDistributedMask-Eval");
  return MRCNN_OK;
}

// ------------------------------------------------------------------------------
// NCCL (dlopen: no link-time dependency; only multi-GPU callers need it)
// ------------------------------------------------------------------------------
struct NcclId { char internal[128]; };
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load(mrcnn_ctx* ctx) {
  if (g_nccl.h) return MRCNN_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  void* h = nullptr;
  for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  if (!h) return mrcnn_fail(ctx, MRCNN_ENCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy)
    return mrcnn_fail(ctx, MRCNN_ENCCL, "libnccl.so.2 lacks a required symbol");
  g_nccl.h = h;
  return MRCNN_OK;
}

static int nccl_fail(mrcnn_ctx* ctx, const char* what, int r) {
  return mrcnn_fail(ctx, MRCNN_ENCCL, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error"));
}

void comm_destroy(mrcnn_ctx* ctx) {
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}

// packs / unpacks the all-gather payload: one row per image = [D*6 detections | D*S*S masks]
__global__ void pack_rows_kernel(const float* __restrict__ det, const float* __restrict__ masks, int n_det, int n_mask,
                                 float* __restrict__ packed) {
  const int img = blockIdx.y;
  const int row = n_det + n_mask;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row; i += gridDim.x * blockDim.x)
    packed[(size_t)img * row + i] = i < n_det ? det[(size_t)img * n_det + i] : masks[(size_t)img * n_mask + (i - n_det)];
}
__global__ void unpack_rows_kernel(const float* __restrict__ packed, int n_det, int n_mask, float* __restrict__ det,
                                   float* __restrict__ masks) {
  const int img = blockIdx.y;
  const int row = n_det + n_mask;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row; i += gridDim.x * blockDim.x) {
    const float v = packed[(size_t)img * row + i];
    if (i < n_det) det[(size_t)img * n_det + i] = v; else masks[(size_t)img * n_mask + (i - n_det)] = v;
  }
}

static int allgather_reserve(mrcnn_ctx* ctx, int batch_local) {
  const mrcnn_config& cfg = ctx->cfg;
  const int D = cfg.max_detections, S = 2 * cfg.pool_size_mask;
  const int n_det = D * 6, n_mask = D * S * S, row = n_det + n_mask;
  const int total = batch_local * ctx->nranks;
  void* t;
  TRY(get_buf(ctx, "ag_det", sizeof(float) * n_det * batch_local, &t));
  TRY(get_buf(ctx, "ag_mask", sizeof(float) * (size_t)n_mask * batch_local, &t));
  TRY(get_buf(ctx, "ag_send", sizeof(float) * (size_t)row * batch_local, &t));
  TRY(get_buf(ctx, "ag_recv", sizeof(float) * (size_t)row * total, &t));
  return MRCNN_OK;
}

// predict on this rank's images, then the one exchange step of the path: a single all-gather of the packed rows over
// NVLink.  Device pointers; buffers reserved by allgather_reserve.
static int predict_allgather_device(mrcnn_ctx* ctx, int batch_local, const uint8_t* drgb, float* ddet_all, float* dmask_all,
                                    int parity = 0, cudaStream_t heads = nullptr) {
  DenseModel* m = model_of(ctx);
  const mrcnn_config& cfg = ctx->cfg;
  const int D = cfg.max_detections, S = 2 * cfg.pool_size_mask;
  const int n_det = D * 6, n_mask = D * S * S, row = n_det + n_mask;
  const int total = batch_local * ctx->nranks;
  float* ldet = (float*)m->bufs["ag_det"].p; float* lmask = (float*)m->bufs["ag_mask"].p;
  float* send = (float*)m->bufs["ag_send"].p; float* recv = (float*)m->bufs["ag_recv"].p;
  TRY(predict_device(ctx, batch_local, drgb, ldet, lmask, parity, heads));
  StreamGuard guard(ctx);
  if (heads) ctx->stream = heads;        // the exchange follows the heads section on its stream: under the next batch's backbone
  {
    ProfScope ps(ctx, PROF_GLUE, (double)row * 4.0 * (batch_local * 2 + total * 2));
    pack_rows_kernel<<<dim3(32, batch_local), 256, 0, ctx->stream>>>(ldet, lmask, n_det, n_mask, send);
    MRCNN_LAUNCH_CHECK(ctx);
    int r = g_nccl.AllGather(send, recv, (size_t)row * batch_local, /*ncclFloat*/ 7, ctx->nccl_comm, ctx->stream);
    if (r != 0) return nccl_fail(ctx, "ncclAllGather", r);
    unpack_rows_kernel<<<dim3(32, total), 256, 0, ctx->stream>>>(recv, n_det, n_mask, ddet_all, dmask_all);
    MRCNN_LAUNCH_CHECK(ctx);
  }
  stage_mark(ctx, "AllGather");
  return MRCNN_OK;
}

extern "C" {

MRCNN_API int mrcnn_nccl_unique_id(void* id_out_128) {
  if (!id_out_128) return MRCNN_EINVAL;
  int rc = nccl_load(nullptr);
  if (rc) return rc;
  int r = g_nccl.GetUniqueId(id_out_128);
  return r == 0 ? MRCNN_OK : nccl_fail(nullptr, "ncclGetUniqueId", r);
}

MRCNN_API int mrcnn_comm_init(mrcnn_ctx* ctx, const void* id128, int rank, int nranks) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, id128 && nranks >= 1 && rank >= 0 && rank < nranks, "comm_init: bad argument");
  TRY(nccl_load(ctx));
  cudaSetDevice(ctx->device);
  comm_destroy(ctx);
  NcclId id;
  memcpy(&id, id128, 128);
  int r = g_nccl.CommInitRank(&ctx->nccl_comm, nranks, id, rank);
  if (r != 0) { ctx->nccl_comm = nullptr; return nccl_fail(ctx, "ncclCommInitRank", r); }
  ctx->rank = rank; ctx->nranks = nranks;
  return MRCNN_OK;
}

// ---- layer-level dense entry points ------------------------------------------------------
MRCNN_API int mrcnn_classifier_eval(mrcnn_ctx* ctx, int batch, int64_t R, const float* pooled, float* out) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, pooled && out && batch >= 1 && R >= 1, "classifier_eval: bad argument");
  cudaSetDevice(ctx->device);
  const int P = ctx->cfg.pool_size_classifier;
  const int64_t M = (int64_t)batch * R;
  std::shared_ptr<Graph> g;
  TRY(cls_graph(ctx, M, &g));
  TRY(cls_graph(ctx, M, &g));      // second fetch: the first may have re-allocated shared buffers
  Stager st(ctx);
  int rc = MRCNN_OK;
  const float* dp = (const float*)st.in(pooled, sizeof(float) * M * 256 * P * P, &rc);
  float* dout = (float*)st.out(out, sizeof(float) * 6 * M, &rc);
  if (rc) return mrcnn_fail(ctx, rc, "classifier_eval: staging failed");
  {
    ProfScope ps(ctx, PROF_GLUE, (double)M * 256 * P * P * 6.0);
    chw_f32_to_nhwc_f16_kernel<<<(unsigned)M, 256, 0, ctx->stream>>>(dp, 256, P * P, (__half*)model_of(ctx)->bufs["pooled_cls"].p, nullptr);
    MRCNN_LAUNCH_CHECK(ctx);
  }
  TRY(run_graph(ctx, *g));
  TRY(cls_post(ctx, M, dout, nullptr, nullptr));
  return st.finish();
}

MRCNN_API int mrcnn_mask_eval(mrcnn_ctx* ctx, int batch, int64_t D, const float* pooled, const float* detections, float* out) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, pooled && detections && out && batch >= 1 && D >= 1, "mask_eval: bad argument");
  cudaSetDevice(ctx->device);
  const int P = ctx->cfg.pool_size_mask, S = 2 * P;
  const int64_t M = (int64_t)batch * D;
  std::shared_ptr<Graph> g;
  TRY(mask_graph(ctx, M, &g));
  TRY(mask_graph(ctx, M, &g));
  Stager st(ctx);
  int rc = MRCNN_OK;
  const float* dp = (const float*)st.in(pooled, sizeof(float) * M * 256 * P * P, &rc);
  const float* dd = (const float*)st.in(detections, sizeof(float) * 6 * M, &rc);
  float* dout = (float*)st.out(out, sizeof(float) * M * S * S, &rc);
  if (rc) return mrcnn_fail(ctx, rc, "mask_eval: staging failed");
  DenseModel* m = model_of(ctx);
  {
    ProfScope ps(ctx, PROF_GLUE, (double)M * 256 * P * P * 6.0);
    // removeZeros (TimeDistributedClassifierLayer.swift:116-127, intended reading Q9): a block is valid iff it is not all zero
    chw_f32_to_nhwc_f16_kernel<<<(unsigned)M, 256, 0, ctx->stream>>>(dp, 256, P * P, (__half*)m->bufs["pooled_mask"].p,
                                                                    (int32_t*)m->bufs["mask_valid"].p);
    MRCNN_LAUNCH_CHECK(ctx);
  }
  TRY(run_graph(ctx, *g));
  TRY(mask_tail(ctx, batch, (int)D, dd, dout));
  return st.finish();
}

MRCNN_API int mrcnn_backbone_eval(mrcnn_ctx* ctx, int batch, const uint8_t* rgb, void* const fmaps_out[4], float* probs_out,
                                  float* deltas_out) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, rgb && batch >= 1, "backbone_eval: bad argument");
  cudaSetDevice(ctx->device);
  DenseModel* m = model_of(ctx);
  std::shared_ptr<Graph> g;
  TRY(backbone_graph(ctx, batch, &g));
  TRY(backbone_graph(ctx, batch, &g));
  Stager st(ctx);
  int rc = MRCNN_OK;
  const mrcnn_config& cfg = ctx->cfg;
  const uint8_t* drgb = (const uint8_t*)st.in(rgb, (size_t)batch * cfg.image_h * cfg.image_w * 3, &rc);
  if (rc) return mrcnn_fail(ctx, rc, "backbone_eval: staging failed");
  m->rgb = drgb;
  TRY(run_graph(ctx, *g));
  const char* pn[4] = {"p2", "p3", "p4", "p5"};
  for (int l = 0; l < 4 && fmaps_out; ++l)
    if (fmaps_out[l])
      MRCNN_CUDA_TRY(ctx, cudaMemcpyAsync(fmaps_out[l], m->bufs[pn[l]].p, (size_t)batch * m->lvl_h[l] * m->lvl_w[l] * 256 * 2,
                                          cudaMemcpyDefault, ctx->stream));
  if (probs_out) MRCNN_CUDA_TRY(ctx, cudaMemcpyAsync(probs_out, m->bufs["rpn_probs"].p, (size_t)batch * m->n_anchors * 8, cudaMemcpyDefault, ctx->stream));
  if (deltas_out) MRCNN_CUDA_TRY(ctx, cudaMemcpyAsync(deltas_out, m->bufs["rpn_deltas"].p, (size_t)batch * m->n_anchors * 16, cudaMemcpyDefault, ctx->stream));
  int frc = st.finish();
  if (frc) return frc;
  MRCNN_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // outputs may be host memory
  return MRCNN_OK;
}

// ---- the MaskRCNN model's prediction ---------------------------------------------------------
MRCNN_API int mrcnn_predict(mrcnn_ctx* ctx, int batch, const uint8_t* rgb, float* detections, float* masks) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, rgb && detections && masks && batch >= 1, "predict: bad argument");
  MRCNN_REQUIRE(ctx, mrcnn_predict_in_flight(ctx) == 0, "predict: streamed batches are in flight; call mrcnn_predict_wait first");
  cudaSetDevice(ctx->device);
  const mrcnn_config& cfg = ctx->cfg;
  const int D = cfg.max_detections, S = 2 * cfg.pool_size_mask;
  Stager st(ctx);
  int rc = MRCNN_OK;
  const uint8_t* drgb = (const uint8_t*)st.in(rgb, (size_t)batch * cfg.image_h * cfg.image_w * 3, &rc);
  float* ddet = (float*)st.out(detections, sizeof(float) * 6 * D * batch, &rc);
  float* dmask = (float*)st.out(masks, sizeof(float) * (size_t)S * S * D * batch, &rc);
  if (rc) return mrcnn_fail(ctx, rc, "predict: staging failed");
  TRY(predict_device(ctx, batch, drgb, ddet, dmask));
  return st.finish();   // stage times are collected lazily by mrcnn_last_stage_times
}

MRCNN_API int mrcnn_predict_allgather(mrcnn_ctx* ctx, int batch_local, const uint8_t* rgb, float* detections_all, float* masks_all) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, rgb && detections_all && masks_all && batch_local >= 1, "predict_allgather: bad argument");
  MRCNN_REQUIRE(ctx, ctx->nccl_comm, "predict_allgather: call mrcnn_comm_init first");
  MRCNN_REQUIRE(ctx, mrcnn_predict_in_flight(ctx) == 0, "predict_allgather: streamed batches are in flight; call mrcnn_predict_wait first");
  cudaSetDevice(ctx->device);
  const mrcnn_config& cfg = ctx->cfg;
  const int D = cfg.max_detections, S = 2 * cfg.pool_size_mask;
  const int total = batch_local * ctx->nranks;
  TRY(allgather_reserve(ctx, batch_local));
  Stager st(ctx);
  int rc = MRCNN_OK;
  const uint8_t* drgb = (const uint8_t*)st.in(rgb, (size_t)batch_local * cfg.image_h * cfg.image_w * 3, &rc);
  float* ddet = (float*)st.out(detections_all, sizeof(float) * 6 * D * total, &rc);
  float* dmask = (float*)st.out(masks_all, sizeof(float) * (size_t)S * S * D * total, &rc);
  if (rc) return mrcnn_fail(ctx, rc, "predict_allgather: staging failed");
  TRY(predict_allgather_device(ctx, batch_local, drgb, ddet, dmask));
  return st.finish();
}

// ---- streaming prediction: at most two batches in flight ---------------------------------------
MRCNN_API int mrcnn_predict_submit(mrcnn_ctx* ctx, int batch, const uint8_t* rgb, float* detections, float* masks, int flags) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, rgb && detections && masks && batch >= 1, "predict_submit: bad argument");
  MRCNN_REQUIRE(ctx, (flags & ~MRCNN_SUBMIT_ALLGATHER) == 0, "predict_submit: unknown flag");
  const bool gather = (flags & MRCNN_SUBMIT_ALLGATHER) != 0;
  if (gather) MRCNN_REQUIRE(ctx, ctx->nccl_comm, "predict_submit: MRCNN_SUBMIT_ALLGATHER needs mrcnn_comm_init first");
  cudaSetDevice(ctx->device);
  DenseModel* m = model_of(ctx);
  MRCNN_REQUIRE(ctx, m->submitted - m->completed < 2, "predict_submit: two batches are already in flight; call mrcnn_predict_wait first");
  const mrcnn_config& cfg = ctx->cfg;
  const int D = cfg.max_detections, S = 2 * cfg.pool_size_mask;
  const int total = gather ? batch * ctx->nranks : batch;
  const int k = (int)(m->submitted & 1);
  DenseModel::StreamSlot& sl = m->slot[k];
  if (!m->copy_stream) MRCNN_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  if (!m->copy_out_stream) MRCNN_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&m->copy_out_stream, cudaStreamNonBlocking));
  if (!sl.done) {
    MRCNN_CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
    MRCNN_CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl.computed, cudaEventDisableTiming));
    MRCNN_CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  }
  // Per-slot device mirrors of host arguments.  A slot is only reused after the batch that used it was waited for
  // (the two-in-flight rule above), so nothing on the device still reads or writes them here.
  const size_t in_bytes = (size_t)batch * cfg.image_h * cfg.image_w * 3;
  const size_t det_bytes = sizeof(float) * 6 * D * total, mask_bytes = sizeof(float) * (size_t)S * S * D * total;
  const char* nm_in[2] = {"stream_in0", "stream_in1"};
  const char* nm_det[2] = {"stream_det0", "stream_det1"};
  const char* nm_mask[2] = {"stream_mask0", "stream_mask1"};
  const bool in_host = !Stager::is_device_ptr(rgb), det_host = !Stager::is_device_ptr(detections),
             mask_host = !Stager::is_device_ptr(masks);
  // every buffer the graphs and this call need exists before anything is enqueued (growing one synchronises)
  {
    const int MBs = cfg.max_batch > batch ? cfg.max_batch : batch;
    void* tmp;
    if (in_host) TRY(get_buf(ctx, nm_in[k], (size_t)MBs * cfg.image_h * cfg.image_w * 3, &tmp));
    if (det_host) TRY(get_buf(ctx, nm_det[k], det_bytes, &tmp));
    if (mask_host) TRY(get_buf(ctx, nm_mask[k], mask_bytes, &tmp));
    if (gather) TRY(allgather_reserve(ctx, batch));
  }
  const uint8_t* drgb = rgb;
  if (in_host) {
    // the copy of this batch overlaps the compute of the previous one: own stream, then an event edge
    drgb = (const uint8_t*)m->bufs[nm_in[k]].p;
    MRCNN_CUDA_TRY(ctx, cudaMemcpyAsync((void*)drgb, rgb, in_bytes, cudaMemcpyHostToDevice, m->copy_stream));
    MRCNN_CUDA_TRY(ctx, cudaEventRecord(sl.h2d_done, m->copy_stream));
    MRCNN_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, sl.h2d_done, 0));
  }
  float* ddet = det_host ? (float*)m->bufs[nm_det[k]].p : detections;
  float* dmask = mask_host ? (float*)m->bufs[nm_mask[k]].p : masks;
  // two-stage software pipeline: the backbone of this batch runs on the context's stream right behind the previous
  // batch's backbone, everything after it on the heads stream behind the previous batch's heads
  // (MRCNN_STREAM_HEADS=0: one stream, as the blocking call)
  static int use_heads = -1;
  if (use_heads < 0) { const char* e = getenv("MRCNN_STREAM_HEADS"); use_heads = e ? atoi(e) != 0 : 1; }
  cudaStream_t heads = nullptr;
  if (use_heads) {
    if (!m->heads_stream) MRCNN_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&m->heads_stream, cudaStreamNonBlocking));
    heads = m->heads_stream;
  }
  cudaStream_t tail = heads ? heads : ctx->stream;
  if (gather) TRY(predict_allgather_device(ctx, batch, drgb, ddet, dmask, k, heads));
  else TRY(predict_device(ctx, batch, drgb, ddet, dmask, k, heads));
  if (det_host || mask_host) {
    // results leave on their own stream, under the compute of the next batch (which writes the other slot's mirrors)
    MRCNN_CUDA_TRY(ctx, cudaEventRecord(sl.computed, tail));
    MRCNN_CUDA_TRY(ctx, cudaStreamWaitEvent(m->copy_out_stream, sl.computed, 0));
    if (det_host) MRCNN_CUDA_TRY(ctx, cudaMemcpyAsync(detections, ddet, det_bytes, cudaMemcpyDeviceToHost, m->copy_out_stream));
    if (mask_host) MRCNN_CUDA_TRY(ctx, cudaMemcpyAsync(masks, dmask, mask_bytes, cudaMemcpyDeviceToHost, m->copy_out_stream));
    MRCNN_CUDA_TRY(ctx, cudaEventRecord(sl.done, m->copy_out_stream));
  } else {
    MRCNN_CUDA_TRY(ctx, cudaEventRecord(sl.done, tail));
  }
  m->submitted++;
  return MRCNN_OK;
}

MRCNN_API int mrcnn_predict_wait(mrcnn_ctx* ctx) {
  if (!ctx) return MRCNN_EINVAL;
  cudaSetDevice(ctx->device);
  DenseModel* m = model_of(ctx);
  MRCNN_REQUIRE(ctx, m->completed < m->submitted, "predict_wait: nothing in flight");
  const int k = (int)(m->completed & 1);
  m->completed++;                   // the slot is released even when the wait reports an error
  MRCNN_CUDA_TRY(ctx, cudaEventSynchronize(m->slot[k].done));
  return MRCNN_OK;
}

MRCNN_API int mrcnn_predict_in_flight(const mrcnn_ctx* ctx) {
  if (!ctx || !ctx->dense) return 0;
  return (int)(ctx->dense->submitted - ctx->dense->completed);
}

}  // extern "C"

// mrcnn_last_stage_times (api.cu) reads ctx->stage_ms; refresh it from the events of the last predict.
void dense_collect_stage_times(mrcnn_ctx* ctx) { if (ctx->dense) stage_collect(ctx); }
