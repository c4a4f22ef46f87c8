#include "dense.h"
int dense_load_weights(mrcnn_ctx* ctx, int, const void*, size_t) { return mrcnn_fail(ctx, MRCNN_ESTATE, "dense model not built"); }
void dense_destroy(mrcnn_ctx*) {}
void comm_destroy(mrcnn_ctx*) {}
extern "C" {
MRCNN_API int mrcnn_classifier_eval(mrcnn_ctx* ctx, int, int64_t, const float*, float*) { return mrcnn_fail(ctx, MRCNN_ESTATE, "nyi"); }
MRCNN_API int mrcnn_mask_eval(mrcnn_ctx* ctx, int, int64_t, const float*, const float*, float*) { return mrcnn_fail(ctx, MRCNN_ESTATE, "nyi"); }
MRCNN_API int mrcnn_predict(mrcnn_ctx* ctx, int, const uint8_t*, float*, float*) { return mrcnn_fail(ctx, MRCNN_ESTATE, "nyi"); }
MRCNN_API int mrcnn_nccl_unique_id(void*) { return MRCNN_ESTATE; }
MRCNN_API int mrcnn_comm_init(mrcnn_ctx* ctx, const void*, int, int) { return mrcnn_fail(ctx, MRCNN_ESTATE, "nyi"); }
MRCNN_API int mrcnn_predict_allgather(mrcnn_ctx* ctx, int, const uint8_t*, float*, float*) { return mrcnn_fail(ctx, MRCNN_ESTATE, "nyi"); }
MRCNN_API int mrcnn_backbone_eval(mrcnn_ctx* ctx, int, const uint8_t*, void* const*, float*, float*) { return mrcnn_fail(ctx, MRCNN_ESTATE, "nyi"); }
}
