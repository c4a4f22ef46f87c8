// letterbox.cu -- the step in front of the path: Vision's `.scaleFit`
// (EvaluateCommand.swift:157, ViewController.swift:42) delivers the image scaled to fit the model's square input with
// centred padding.  Geometry follows the reference's own letterboxing code (DetectionRenderer.swift:63-75: scale factor
// chosen by `fitsHorizontally`, padding split evenly); the resampling filter of Vision is closed, so bilinear with
// half-pixel centres is used (PARITY UNPINNED for the filter).  Arithmetic in fp64, op by op, identical to
// oracle/oracle.c orc_letterbox -> bit-exact.
#include "common.cuh"

struct LetterboxGeom { double scale, new_w, new_h, pad_x, pad_y; };

static LetterboxGeom letterbox_geom(int src_h, int src_w, int dst_h, int dst_w) {
  LetterboxGeom g;
  const double hs = (double)dst_w / (double)src_w, vs = (double)dst_h / (double)src_h;   // DetectionRenderer.swift:63-64
  const bool fits_h = (double)src_h * hs <= (double)dst_h;                                // :66
  g.scale = fits_h ? hs : vs;                                                             // :68
  g.new_w = (double)src_w * g.scale; g.new_h = (double)src_h * g.scale;                   // :70
  g.pad_x = ((double)dst_w - g.new_w) / 2.0; g.pad_y = ((double)dst_h - g.new_h) / 2.0;   // :72-75
  return g;
}

__global__ void letterbox_kernel(const uint8_t* __restrict__ src, int src_h, int src_w, int dst_h, int dst_w,
                                 LetterboxGeom g, uint8_t* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)dst_h * dst_w) return;
  const int X = (int)(i % dst_w), Y = (int)(i / dst_w);
  const double cx = __dsub_rn(__dadd_rn((double)X, 0.5), g.pad_x), cy = __dsub_rn(__dadd_rn((double)Y, 0.5), g.pad_y);
  uint8_t* o = dst + i * 3;
  if (!(cx >= 0.0 && cx < g.new_w && cy >= 0.0 && cy < g.new_h)) { o[0] = o[1] = o[2] = 0; return; }
  double sx = __dsub_rn(__ddiv_rn(cx, g.scale), 0.5), sy = __dsub_rn(__ddiv_rn(cy, g.scale), 0.5);
  sx = sx < 0.0 ? 0.0 : (sx > (double)(src_w - 1) ? (double)(src_w - 1) : sx);
  sy = sy < 0.0 ? 0.0 : (sy > (double)(src_h - 1) ? (double)(src_h - 1) : sy);
  const int x0 = (int)floor(sx), y0 = (int)floor(sy);
  const int x1 = x0 + 1 < src_w ? x0 + 1 : x0, y1 = y0 + 1 < src_h ? y0 + 1 : y0;
  const double fx = __dsub_rn(sx, (double)x0), fy = __dsub_rn(sy, (double)y0);
  for (int c = 0; c < 3; ++c) {
    const double tl = src[((int64_t)y0 * src_w + x0) * 3 + c], tr = src[((int64_t)y0 * src_w + x1) * 3 + c];
    const double bl = src[((int64_t)y1 * src_w + x0) * 3 + c], br = src[((int64_t)y1 * src_w + x1) * 3 + c];
    const double top = __dadd_rn(tl, __dmul_rn(__dsub_rn(tr, tl), fx));
    const double bot = __dadd_rn(bl, __dmul_rn(__dsub_rn(br, bl), fx));
    const double v = __dadd_rn(top, __dmul_rn(__dsub_rn(bot, top), fy));
    o[c] = (uint8_t)__dadd_rn(v, 0.5);
  }
}

extern "C" {

MRCNN_API int mrcnn_letterbox_geometry(int src_h, int src_w, int dst_h, int dst_w, double out5[5]) {
  if (!out5 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1) return MRCNN_EINVAL;
  const LetterboxGeom g = letterbox_geom(src_h, src_w, dst_h, dst_w);
  out5[0] = g.scale; out5[1] = g.new_w; out5[2] = g.new_h; out5[3] = g.pad_x; out5[4] = g.pad_y;
  return MRCNN_OK;
}

MRCNN_API int mrcnn_letterbox_eval(mrcnn_ctx* ctx, const uint8_t* src, int src_h, int src_w, uint8_t* dst) {
  if (!ctx) return MRCNN_EINVAL;
  MRCNN_REQUIRE(ctx, src && dst && src_h >= 1 && src_w >= 1, "letterbox: bad argument");
  cudaSetDevice(ctx->device);
  const int dh = ctx->cfg.image_h, dw = ctx->cfg.image_w;
  Stager st(ctx);
  int rc = MRCNN_OK;
  const uint8_t* ds = (const uint8_t*)st.in(src, (size_t)src_h * src_w * 3, &rc);
  uint8_t* dd = (uint8_t*)st.out(dst, (size_t)dh * dw * 3, &rc);
  if (rc) return mrcnn_fail(ctx, rc, "letterbox: staging failed");
  {
    ProfScope ps(ctx, PROF_GLUE, 3.0 * ((double)src_h * src_w + (double)dh * dw));
    letterbox_kernel<<<ceil_div((int64_t)dh * dw, 256), 256, 0, ctx->stream>>>(ds, src_h, src_w, dh, dw, letterbox_geom(src_h, src_w, dh, dw), dd);
    MRCNN_LAUNCH_CHECK(ctx);
  }
  return st.finish();
}

// Boxes (y1,x1,y2,x2) normalised to the letter-boxed model frame -> normalised to the source image (host arithmetic;
// the inverse of the mapping above).  boxes / out: n x row_stride floats, first 4 of each row are the box.
MRCNN_API int mrcnn_unletterbox_boxes(int src_h, int src_w, int dst_h, int dst_w, const float* boxes, int64_t n, int row_stride,
                                      float* out) {
  if (!boxes || !out || n < 0 || row_stride < 4 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1) return MRCNN_EINVAL;
  const LetterboxGeom g = letterbox_geom(src_h, src_w, dst_h, dst_w);
  for (int64_t i = 0; i < n; ++i) {
    const float* b = boxes + i * row_stride;
    float* o = out + i * row_stride;
    for (int k = 4; k < row_stride; ++k) o[k] = b[k];
    o[0] = (float)(((double)b[0] * dst_h - g.pad_y) / g.new_h);
    o[1] = (float)(((double)b[1] * dst_w - g.pad_x) / g.new_w);
    o[2] = (float)(((double)b[2] * dst_h - g.pad_y) / g.new_h);
    o[3] = (float)(((double)b[3] * dst_w - g.pad_x) / g.new_w);
  }
  return MRCNN_OK;
}

}  // extern "C"
