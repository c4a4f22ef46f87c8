/* Minimal C caller of libmaskrcnn_cuda.so: the same calls the Swift shim (swift/Sources/MaskRCNNCuda/MaskRCNN.swift) and
 * the Python mirror make.  Streams `n_batches` batches of synthetic 1024x1024 images through mrcnn_predict_submit /
 * mrcnn_predict_wait and decodes the detections of the last one.
 *
 *   gcc -std=c99 -Iinclude examples/predict.c -Lmask-rcnn-coreml_b200 -lmaskrcnn_cuda -o predict_example
 *   LD_LIBRARY_PATH=mask-rcnn-coreml_b200 ./predict_example products/ 8 4
 * products/ holds anchors.bin, MaskRCNN.mrcnnw, Classifier.mrcnnw, Mask.mrcnnw (weights.write_products or
 * tools/import_mlmodel.py). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "maskrcnn_cuda.h"

static int fail(mrcnn_ctx* ctx, const char* what, int status) {
  fprintf(stderr, "%s failed (%d): %s\n", what, status, mrcnn_last_error(ctx));
  return 1;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s <products dir> [batch] [n_batches]\n", argv[0]); return 2; }
  const int batch = argc > 2 ? atoi(argv[2]) : 8, n_batches = argc > 3 ? atoi(argv[3]) : 4;
  char anchors[1024], mainp[1024], clsp[1024], maskp[1024];
  snprintf(anchors, sizeof anchors, "%s/anchors.bin", argv[1]);
  snprintf(mainp, sizeof mainp, "%s/MaskRCNN.mrcnnw", argv[1]);
  snprintf(clsp, sizeof clsp, "%s/Classifier.mrcnnw", argv[1]);
  snprintf(maskp, sizeof maskp, "%s/Mask.mrcnnw", argv[1]);

  mrcnn_config cfg;
  mrcnn_config_default(&cfg);                 /* 1024x1024, resnet101, 81 classes, 6000 -> 1000 proposals, 100 detections */
  cfg.max_batch = batch;
  cfg.anchors_path = anchors; cfg.main_model_path = mainp; cfg.classifier_model_path = clsp; cfg.mask_model_path = maskp;
  mrcnn_ctx* ctx = NULL;
  int st = mrcnn_create(&cfg, &ctx);
  if (st) return fail(NULL, "mrcnn_create", st);
  printf("%s\n", mrcnn_version());

  const size_t img_bytes = (size_t)batch * cfg.image_h * cfg.image_w * 3;
  const int D = cfg.max_detections, S = 2 * cfg.pool_size_mask;
  uint8_t* img[2]; float* det[2]; float* msk[2];
  for (int k = 0; k < 2; ++k) {                /* two batches in flight: two sets of host buffers */
    img[k] = (uint8_t*)malloc(img_bytes);
    det[k] = (float*)malloc(sizeof(float) * batch * D * 6);
    msk[k] = (float*)malloc(sizeof(float) * (size_t)batch * D * S * S);
    if (!img[k] || !det[k] || !msk[k]) return 3;
    for (size_t i = 0; i < img_bytes; ++i) img[k][i] = (uint8_t)((i * 2654435761u + (unsigned)k) >> 24);
  }
  for (int i = 0; i < n_batches; ++i) {
    st = mrcnn_predict_submit(ctx, batch, img[i & 1], det[i & 1], msk[i & 1], 0);
    if (st) return fail(ctx, "mrcnn_predict_submit", st);
    if (i >= 1 && (st = mrcnn_predict_wait(ctx)) != 0) return fail(ctx, "mrcnn_predict_wait", st);
  }
  if ((st = mrcnn_predict_wait(ctx)) != 0) return fail(ctx, "mrcnn_predict_wait", st);

  /* Detection.detectionsFromFeatureValue (Detection.swift:23-62) on the last batch */
  const int last = (n_batches - 1) & 1;
  int32_t* count = (int32_t*)calloc(batch, sizeof(int32_t));
  int32_t* index = (int32_t*)calloc((size_t)batch * D, sizeof(int32_t));
  int32_t* cls = (int32_t*)calloc((size_t)batch * D, sizeof(int32_t));
  double* bbox = (double*)calloc((size_t)batch * D * 4, sizeof(double));
  double* score = (double*)calloc((size_t)batch * D, sizeof(double));
  st = mrcnn_detections_decode(ctx, batch, det[last], msk[last], count, index, bbox, cls, score, NULL);
  if (st) return fail(ctx, "mrcnn_detections_decode", st);
  for (int b = 0; b < batch; ++b) printf("image %d: %d detections\n", b, (int)count[b]);

  const char* names[16]; float ms[16];
  const int n = mrcnn_last_stage_times(ctx, 16, names, ms);
  for (int i = 0; i < n; ++i) printf("  %-40s %.3f ms\n", names[i], ms[i]);
  mrcnn_destroy(ctx);
  return 0;
}
