// roialign.cu -- PyramidROIAlignLayer.evaluate (PyramidROIAlignLayer.swift:79-181).
//
//   roi_level_kernel      roisToInputItems (:351-396): FPN level per roi in fp64,
//                         round half away from zero, clamp 2..5; NaN/inf -> padding
//   roialign_chw_kernel   boundary layout (reference layout): maps CHW fp32,
//                         output (R,C,P,P) fp32 (copyOutput :245-274 order)
//   roialign_nhwc_kernel  internal layout of the fused pipeline: maps NHWC fp16,
//                         output (R,P,P,C) fp16 (the K-major operand of the head GEMMs)
// Sampling = MPSNNCropAndResizeBilinear (:212-223) restated as TensorFlow
// crop_and_resize (bilinear, extrapolation 0) in fp32, every op rounded
// individually so the result is bit-identical to oracle/oracle.c.
// Every output block is written (fixes Q5: the reference drops the last group).
#include "common.cuh"
#include "exact_math.cuh"

__global__ void roi_level_kernel(const float* __restrict__ rois, int roi_stride, int64_t total,
                                 double ratio, int32_t* __restrict__ level) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float* r = rois + i * roi_stride;
  double y1 = r[0], x1 = r[1], y2 = r[2], x2 = r[3];
  double w = __dsub_rn(x2, x1), h = __dsub_rn(y2, y1);
  double lf = __dadd_rn(log2(__ddiv_rn(sqrt(__dmul_rn(w, h)), ratio)), 4.0);   // :373
  int lv;
  if (isnan(lf) || isinf(lf)) lv = -1;                                          // :374 -> padding
  else {
    double rl = round(lf);                      // half away from zero (Q13)
    rl = rl < 2.0 ? 2.0 : (rl > 5.0 ? 5.0 : rl);
    lv = (int)rl;                                                               // :376
  }
  level[i] = lv;
}

struct PyramidF32 { const float* p[4]; int h[4]; int w[4]; };
struct PyramidF16 { const __half* p[4]; int h[4]; int w[4]; };

struct SampleAxis { float lerp; int lo, hi; bool ok; };

// in = a1*(D-1) + idx * ((a2-a1)*(D-1)/(P-1))   (TF crop_and_resize, fp32)
__device__ __forceinline__ SampleAxis sample_axis(float a1, float a2, int D, int P, int idx) {
  const float dm1 = (float)(D - 1);
  float in;
  if (P > 1) {
    float scale = __fdiv_rn(__fmul_rn(__fsub_rn(a2, a1), dm1), (float)(P - 1));
    in = __fadd_rn(__fmul_rn(a1, dm1), __fmul_rn((float)idx, scale));
  } else {
    in = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(a1, a2)), dm1);
  }
  SampleAxis s;
  s.ok = !(in < 0.0f || in > dm1);
  float f = floorf(in), c = ceilf(in);
  s.lo = (int)f; s.hi = (int)c;
  s.lerp = __fsub_rn(in, f);
  return s;
}

__device__ __forceinline__ float bilerp(float tl, float tr, float bl, float br, float lx, float ly) {
  float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
  float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
  return __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
}

// v0 (gather): grid (chunks of C*P*P outputs, R, batch). One thread per output
// element; consecutive threads walk (c, py, px) so writes are fully coalesced.
__global__ void __launch_bounds__(256)
roialign_chw_kernel(const float* __restrict__ rois, int roi_stride, int R, PyramidF32 pyr, int C, int P,
                    const int32_t* __restrict__ level, float* __restrict__ out) {
  const int img = blockIdx.z, r = blockIdx.y;
  const int64_t ri = (int64_t)img * R + r;
  const int PP = P * P;
  const int blk = C * PP;
  float* o = out + ri * blk;
  const int lv = level[ri];
  const int e0 = blockIdx.x * (blockDim.x * 4);
  if (lv < 0) {
    for (int k = 0; k < 4; ++k) { int e = e0 + k * blockDim.x + threadIdx.x; if (e < blk) o[e] = 0.0f; }
    return;
  }
  const int m = lv - 2;
  const int H = pyr.h[m], W = pyr.w[m];
  const float* fm = pyr.p[m] + (size_t)img * C * H * W;
  const float* rr = rois + ri * roi_stride;
  const float y1 = rr[0], x1 = rr[1], y2 = rr[2], x2 = rr[3];
  #pragma unroll
  for (int k = 0; k < 4; ++k) {
    int e = e0 + k * blockDim.x + threadIdx.x;
    if (e >= blk) break;
    int c = e / PP;
    int rem = e - c * PP;
    int py = rem / P, px = rem - py * P;
    SampleAxis sy = sample_axis(y1, y2, H, P, py);
    SampleAxis sx = sample_axis(x1, x2, W, P, px);
    float v = 0.0f;
    if (sy.ok && sx.ok) {
      const float* pl = fm + (size_t)c * H * W;
      float tl = __ldg(pl + (size_t)sy.lo * W + sx.lo), tr = __ldg(pl + (size_t)sy.lo * W + sx.hi);
      float bl = __ldg(pl + (size_t)sy.hi * W + sx.lo), br = __ldg(pl + (size_t)sy.hi * W + sx.hi);
      v = bilerp(tl, tr, bl, br, sx.lerp, sy.lerp);
    }
    o[e] = v;
  }
}

// NHWC fp16: one CTA per roi.  The 2P axis samples (tap rows / columns and lerp weights) are computed once per
// CTA into shared memory; then one warp per output sample, 8 channels (16 B) per lane per step, so every tap is a
// contiguous C*2-byte run.  Two samples are in flight per warp iteration (8 independent 16-byte loads per lane).
__device__ __forceinline__ uint32_t bilerp2(uint32_t a, uint32_t b, uint32_t c, uint32_t d, float lx, float ly) {
  const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&a)), fb = __half22float2(*reinterpret_cast<const __half2*>(&b));
  const float2 fc = __half22float2(*reinterpret_cast<const __half2*>(&c)), fd = __half22float2(*reinterpret_cast<const __half2*>(&d));
  const __half2 r = __floats2half2_rn(bilerp(fa.x, fb.x, fc.x, fd.x, lx, ly), bilerp(fa.y, fb.y, fc.y, fd.y, lx, ly));
  return *reinterpret_cast<const uint32_t*>(&r);
}

__device__ __forceinline__ uint4 bilerp8(uint4 a, uint4 b, uint4 c, uint4 d, float lx, float ly) {
  return make_uint4(bilerp2(a.x, b.x, c.x, d.x, lx, ly), bilerp2(a.y, b.y, c.y, d.y, lx, ly),
                    bilerp2(a.z, b.z, c.z, d.z, lx, ly), bilerp2(a.w, b.w, c.w, d.w, lx, ly));
}

__global__ void __launch_bounds__(256, 3)
roialign_nhwc_kernel(const float* __restrict__ rois, int roi_stride, int R, PyramidF16 pyr, int C, int P,
                     const int32_t* __restrict__ level, __half* __restrict__ out) {
  __shared__ int s_lo[2][64], s_hi[2][64];        // [0] = y axis, [1] = x axis; lo = -1 marks an out-of-range sample
  __shared__ float s_lerp[2][64];
  const int img = blockIdx.y, r = blockIdx.x;
  const int64_t ri = (int64_t)img * R + r;
  const int PP = P * P;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __half* o = out + ri * (int64_t)PP * C;
  const int lv = level[ri];
  const int cvec = C >> 3;   // uint4 (8 halves) per pixel
  if (lv < 0) {
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int e = threadIdx.x; e < PP * cvec; e += blockDim.x) reinterpret_cast<uint4*>(o)[e] = z;
    return;
  }
  const int m = lv - 2;
  const int H = pyr.h[m], W = pyr.w[m];
  const __half* fm = pyr.p[m] + (size_t)img * H * W * C;
  if (threadIdx.x < 128) {
    const int axis = threadIdx.x >> 6, i = threadIdx.x & 63;
    if (i < P) {
      const float* rr = rois + ri * roi_stride;
      const SampleAxis sa = axis == 0 ? sample_axis(rr[0], rr[2], H, P, i) : sample_axis(rr[1], rr[3], W, P, i);
      s_lo[axis][i] = sa.ok ? sa.lo : -1; s_hi[axis][i] = sa.hi; s_lerp[axis][i] = sa.lerp;
    }
  }
  __syncthreads();
  for (int s0 = wid; s0 < PP; s0 += 2 * nw) {
    const int s1 = s0 + nw;
    const bool has1 = s1 < PP;
    const int py0 = s0 / P, px0 = s0 - py0 * P;
    const int py1 = has1 ? s1 / P : py0, px1 = has1 ? s1 - py1 * P : px0;
    const int yl0 = s_lo[0][py0], yh0 = s_hi[0][py0], xl0 = s_lo[1][px0], xh0 = s_hi[1][px0];
    const int yl1 = s_lo[0][py1], yh1 = s_hi[0][py1], xl1 = s_lo[1][px1], xh1 = s_hi[1][px1];
    const bool ok0 = (yl0 >= 0) && (xl0 >= 0), ok1 = has1 && (yl1 >= 0) && (xl1 >= 0);
    for (int v = lane; v < cvec; v += 32) {
      const uint4 z = make_uint4(0, 0, 0, 0);
      uint4 a0 = z, b0 = z, c0 = z, d0 = z, a1 = z, b1 = z, c1 = z, d1 = z;
      if (ok0) {
        a0 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yl0 * W + xl0) * C) + v);
        b0 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yl0 * W + xh0) * C) + v);
        c0 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yh0 * W + xl0) * C) + v);
        d0 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yh0 * W + xh0) * C) + v);
      }
      if (ok1) {
        a1 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yl1 * W + xl1) * C) + v);
        b1 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yl1 * W + xh1) * C) + v);
        c1 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yh1 * W + xl1) * C) + v);
        d1 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yh1 * W + xh1) * C) + v);
      }
      // out-of-range samples are exactly zero (extrapolation value), in-range ones follow the oracle's op order
      reinterpret_cast<uint4*>(o + (size_t)s0 * C)[v] = ok0 ? bilerp8(a0, b0, c0, d0, s_lerp[1][px0], s_lerp[0][py0]) : z;
      if (has1)
        reinterpret_cast<uint4*>(o + (size_t)s1 * C)[v] = ok1 ? bilerp8(a1, b1, c1, d1, s_lerp[1][px1], s_lerp[0][py1]) : z;
    }
  }
}

static int roi_levels(mrcnn_ctx* ctx, int batch, const float* d_rois, int roi_stride, int64_t R,
                      int32_t** d_level) {
  int64_t total = (int64_t)batch * R;
  if (total > ctx->roi_cap) {
    MRCNN_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_roi_level);
    MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_roi_level, sizeof(int32_t) * total));
    ctx->roi_cap = (int)total;
  }
  // PyramidROIAlignLayer.swift:357 ratio = factor / sqrt(W*H)  (Q15: configured size always used)
  double ratio = (double)ctx->cfg.fpn_selection_factor / sqrt((double)ctx->cfg.image_w * (double)ctx->cfg.image_h);
  ProfScope ps(ctx, PROF_GLUE, (double)total * (roi_stride * 4 + 4));
  roi_level_kernel<<<ceil_div(total, 256), 256, 0, ctx->stream>>>(d_rois, roi_stride, total, ratio, ctx->d_roi_level);
  MRCNN_LAUNCH_CHECK(ctx);
  *d_level = ctx->d_roi_level;
  return MRCNN_OK;
}

int roialign_chw_run(mrcnn_ctx* ctx, int batch, const float* d_rois, int roi_stride, int64_t R,
                     const float* const d_fmaps[4], const int32_t hw[8], int64_t C, int P,
                     float* d_out, int32_t* d_level_out) {
  MRCNN_REQUIRE(ctx, batch >= 1 && R >= 1 && R <= 65535, "roialign: bad batch / num_rois");
  MRCNN_REQUIRE(ctx, roi_stride >= 4, "roialign: roi_row_stride must be >= 4");
  MRCNN_REQUIRE(ctx, C >= 1 && P >= 1 && P <= 64, "roialign: bad channels / pool");
  int32_t* lv = nullptr;
  int rc = roi_levels(ctx, batch, d_rois, roi_stride, R, &lv);
  if (rc) return rc;
  PyramidF32 pyr;
  for (int l = 0; l < 4; ++l) {
    pyr.p[l] = d_fmaps[l]; pyr.h[l] = hw[2 * l]; pyr.w[l] = hw[2 * l + 1];
    MRCNN_REQUIRE(ctx, pyr.h[l] >= 1 && pyr.w[l] >= 1, "roialign: bad feature map size");
  }
  const int blk = (int)C * P * P;
  dim3 grid(ceil_div(blk, 1024), (unsigned)R, batch);
  double map_bytes = 0;
  for (int l = 0; l < 4; ++l) map_bytes += 4.0 * (double)C * pyr.h[l] * pyr.w[l];
  // compulsory traffic: every level read once + output written once + rois
  ProfScope ps(ctx, PROF_ROIALIGN, (double)batch * (map_bytes + 4.0 * R * blk + 4.0 * R * roi_stride));
  roialign_chw_kernel<<<grid, 256, 0, ctx->stream>>>(d_rois, roi_stride, (int)R, pyr, (int)C, P, lv, d_out);
  MRCNN_LAUNCH_CHECK(ctx);
  if (d_level_out)
    MRCNN_CUDA_TRY(ctx, cudaMemcpyAsync(d_level_out, lv, sizeof(int32_t) * batch * R, cudaMemcpyDeviceToDevice, ctx->stream));
  return MRCNN_OK;
}

int roialign_nhwc_f16_run(mrcnn_ctx* ctx, int batch, const float* d_rois, int roi_stride, int64_t R,
                          const __half* const d_fmaps[4], const int32_t hw[8], int64_t C, int P,
                          __half* d_out, int32_t* d_level_out) {
  MRCNN_REQUIRE(ctx, batch >= 1 && R >= 1, "roialign: bad batch / num_rois");
  MRCNN_REQUIRE(ctx, roi_stride >= 4, "roialign: roi_row_stride must be >= 4");
  MRCNN_REQUIRE(ctx, C >= 8 && (C % 8) == 0 && P >= 1 && P <= 64, "roialign(nhwc): channels must be a multiple of 8");
  int32_t* lv = nullptr;
  int rc = roi_levels(ctx, batch, d_rois, roi_stride, R, &lv);
  if (rc) return rc;
  PyramidF16 pyr;
  for (int l = 0; l < 4; ++l) { pyr.p[l] = d_fmaps[l]; pyr.h[l] = hw[2 * l]; pyr.w[l] = hw[2 * l + 1]; }
  dim3 grid((unsigned)R, batch);
  double map_bytes = 0;
  for (int l = 0; l < 4; ++l) map_bytes += 2.0 * (double)C * pyr.h[l] * pyr.w[l];
  ProfScope ps(ctx, PROF_ROIALIGN, (double)batch * (map_bytes + 2.0 * R * C * P * P + 4.0 * R * roi_stride));
  roialign_nhwc_kernel<<<grid, 256, 0, ctx->stream>>>(d_rois, roi_stride, (int)R, pyr, (int)C, P, lv, d_out);
  MRCNN_LAUNCH_CHECK(ctx);
  if (d_level_out)
    MRCNN_CUDA_TRY(ctx, cudaMemcpyAsync(d_level_out, lv, sizeof(int32_t) * batch * R, cudaMemcpyDeviceToDevice, ctx->stream));
  return MRCNN_OK;
}
