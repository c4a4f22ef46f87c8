// dense.h -- interface between the C ABI (api.cu) and the dense model
// (dense.cu: tcgen05 implicit-GEMM convolution plans; pipeline.cu: the model).
#pragma once
#include "common.cuh"
#include <string.h>
#include "conv_gemm.cuh"

// One convolution (or GEMM) as the caller sees it.  Activations are NHWC fp16.
struct ConvLaunch {
  const __half* x = nullptr;
  int n = 1, h_in = 1, w_in = 1, cin = 64;
  int ld_in = 0;                 // input pixel stride in elements (0 -> cin)
  // optional explicit A view (overrides the NHWC view): dims fastest-first and byte strides of dims 1..3
  bool custom_view = false;
  uint64_t a_dims[4] = {0, 0, 0, 0};
  uint64_t a_strides[3] = {0, 0, 0};
  const __half* w = nullptr;     // [cout][ntaps*cin] K-major, taps ordered (ky, kx)
  int cout = 64;
  int kh = 1, kw = 1, stride = 1, pad = 0;
  int ntaps_override = 0;        // custom tap list (conv1 space-to-depth view)
  int8_t tap_dx[CG_MAX_TAPS] = {0};
  int8_t tap_dy[CG_MAX_TAPS] = {0};
  const float* bias = nullptr;
  const __half* residual = nullptr;
  int res_mode = 0, res_h = 0, res_w = 0, res_ld = 0;
  int relu = 0, out_f32 = 0, deconv = 0, deconv_c = 0;
  void* out = nullptr;
  int h_out = 0, w_out = 0;      // 0 -> derived from kh/kw/stride/pad
  int ldc = 0;                   // 0 -> round_up(cout, 8)
  int tw = 0, th = 0, bn = 0;    // 0 -> auto
  unsigned long long* trace = nullptr;   // debug event trace buffer (device), see cg::trace_ev
  int ctas = 0;                  // 0 -> auto (env MRCNN_CONV_CTAS overrides), 1 = one CTA per tile, 2 = CTA pairs (cta_group::2)
  int no_tma_epilogue = 0;       // force the direct-store epilogue (tests)
  int no_vgroup = 0;             // one A tile per tap even for 3x3 convolutions
  // mask-head tail fused into the deconv epilogue (needs deconv = 1, deconv_c == bn == 256): out = f32 [n, 2h, 2w]
  int maskdot = 0;
  const int32_t* md_valid = nullptr;
  const int32_t* md_cls = nullptr;
  const __half* md_w = nullptr;
  const float* md_b = nullptr;
  int md_ncls = 0;
  int md_precise = 0;            // keep the deconv activation in fp32 inside the fused tail
  int split_out = 0;             // write fp16 (hi, lo) pairs: out = [n,h,w,2*cout] (see cg::epilogue_split)
};

struct ConvPlan {
  CUtensorMap tmA, tmB, tmC, tmR;
  ConvGemmParams p;
  int bn = 0, grid = 0;
  size_t smem = 0;
  double flops = 0;
};

int conv_plan_build(mrcnn_ctx* ctx, const ConvLaunch& L, ConvPlan* plan);
int conv_plan_run(mrcnn_ctx* ctx, const ConvPlan& plan);

// A bottleneck block's 1x1 expansion (+ residual) and the next block's 1x1 reduction as one launch (conv_fused.cuh)
struct FusedParams {
  int n_img, h, w;             // pixels (both layers: 1x1, stride 1)
  int tw, th, tiles_x, tiles_y;
  int c1, n1, n2;              // channels: A, X, Y
  int relu1, relu2;
  const float* bias1;          // [n1]
  const float* bias2;          // [n2]
  unsigned long long* trace;   // debug event trace (layout of cg::trace_ev: 3 role sections per CTA), or null
};
struct FusedPlan {
  CUtensorMap tmA, tmB1, tmB2, tmR, tmX, tmY;
  FusedParams p;
  int grid = 0, ctas = 1;      // ctas = 2: CTA pairs (tcgen05.mma.cta_group::2), each CTA stages half of every weight tile
  double flops = 0;
};
bool fused_plan_supported(const ConvLaunch& expand, const ConvLaunch& reduce);
int fused_plan_build(mrcnn_ctx* ctx, const ConvLaunch& expand, const ConvLaunch& reduce, FusedPlan* plan);
int fused_plan_run(mrcnn_ctx* ctx, const FusedPlan& plan);

int dense_load_weights(mrcnn_ctx* ctx, int which, const void* blob, size_t bytes);
void dense_destroy(mrcnn_ctx* ctx);
void comm_destroy(mrcnn_ctx* ctx);
void dense_collect_stage_times(mrcnn_ctx* ctx);
