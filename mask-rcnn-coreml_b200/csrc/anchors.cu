// anchors.cu -- on-demand anchor generation (host arithmetic, no device needed).
//
// The reference ships anchors.bin pre-computed by its conversion task (Conversion/task.py:176) and carries the TODO
// "generate the anchors on demand based on image shape, this will save 5mb" (MaskRCNNConfig.swift:14).  The generator
// itself lives in the un-vendored Keras package; this is the Matterport rule it implements:
// scales 32..512 on the pyramid strides 4..64, ratios 0.5 / 1 / 2, anchor stride 1, level-major then y, x, ratio,
// normalised with (box - [0,0,1,1]) / [h-1, w-1, h-1, w-1].  fp64 arithmetic, one rounding to fp32 at the end, in
// the operation order of the Python generator (mask-rcnn-coreml_b200/synth.py) so that both are bit-identical.
#include <math.h>
#include <stdint.h>

#include "../../include/maskrcnn_cuda.h"

namespace {
const int kStrides[5] = {4, 8, 16, 32, 64};
const double kScales[5] = {32, 64, 128, 256, 512};
const double kRatios[3] = {0.5, 1.0, 2.0};
inline int64_t ceil_div_i(int64_t a, int64_t b) { return (a + b - 1) / b; }
}  // namespace

extern "C" {

MRCNN_API int64_t mrcnn_anchor_count(int image_h, int image_w) {
  if (image_h < 2 || image_w < 2) return MRCNN_EINVAL;
  int64_t n = 0;
  for (int l = 0; l < 5; ++l) n += ceil_div_i(image_h, kStrides[l]) * ceil_div_i(image_w, kStrides[l]) * 3;
  return n;
}

MRCNN_API int mrcnn_generate_anchors(int image_h, int image_w, float* anchors_out, int64_t capacity) {
  const int64_t n = mrcnn_anchor_count(image_h, image_w);
  if (n < 0 || !anchors_out || capacity < n) return MRCNN_EINVAL;
  const double sy = (double)(image_h - 1), sx = (double)(image_w - 1);
  float* o = anchors_out;
  for (int l = 0; l < 5; ++l) {
    const int64_t fh = ceil_div_i(image_h, kStrides[l]), fw = ceil_div_i(image_w, kStrides[l]);
    double hh[3], hw[3];
    for (int r = 0; r < 3; ++r) {
      const double root = sqrt(kRatios[r]);
      hh[r] = 0.5 * (kScales[l] / root);
      hw[r] = 0.5 * (kScales[l] * root);
    }
    for (int64_t y = 0; y < fh; ++y) {
      const double cy = (double)y * kStrides[l];
      for (int64_t x = 0; x < fw; ++x) {
        const double cx = (double)x * kStrides[l];
        for (int r = 0; r < 3; ++r) {
          *o++ = (float)(((cy - hh[r]) - 0.0) / sy);
          *o++ = (float)(((cx - hw[r]) - 0.0) / sx);
          *o++ = (float)(((cy + hh[r]) - 1.0) / sy);
          *o++ = (float)(((cx + hw[r]) - 1.0) / sx);
        }
      }
    }
  }
  return MRCNN_OK;
}

}  // extern "C"
