/*
 * maskrcnn_cuda.h -- C ABI of libmaskrcnn_cuda.so, the B200 (sm_100a) drop-in for
 * the Sources/Mask-RCNN-CoreML path of edouardlp/Mask-RCNN-CoreML.
 *
 * The reference exposes this path through Apple's MLCustomLayer protocol
 * (init(parameters:), setWeightData, outputShapes(forInputShapes:),
 * evaluate(inputs:outputs:)) implemented by five @objc classes, plus the
 * Xcode-generated `MaskRCNN` model class on top.  Each entry point below cites
 * the reference interface it replaces (paths relative to
 * Sources/Mask-RCNN-CoreML/ in the reference).  INTEGRATION.md shows the Swift
 * binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - plain C types only; every function returns 0 (MRCNN_OK) or a negative
 *    mrcnn_status; the message is available from mrcnn_last_error().
 *  - every data pointer may be a HOST or a DEVICE pointer (detected with
 *    cudaPointerGetAttributes).  Host buffers are staged through the context's
 *    pinned workspace and the call returns after the stream has been
 *    synchronised; with device pointers the call is stream-ordered and returns
 *    after enqueue.
 *  - layouts at this boundary are the reference's: dense fp32, row-major,
 *    feature maps CHW.  All layer calls take a leading `batch` (images); batch=1
 *    is exactly one reference evaluate().
 *  - the library always writes every output element (the reference relies on
 *    Core ML not clearing buffers; see ProposalLayer.swift:188).
 *  - a context is bound to one CUDA device and one stream and is not
 *    thread-safe; distinct contexts are independent.
 *  - there is no CPU fallback: without a CUDA device mrcnn_create fails.
 */
#ifndef MASKRCNN_CUDA_H_
#define MASKRCNN_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRCNN_ABI_VERSION 2

#if defined(__GNUC__)
#define MRCNN_API __attribute__((visibility("default")))
#else
#define MRCNN_API
#endif

typedef enum mrcnn_status {
  MRCNN_OK = 0,
  MRCNN_EINVAL = -1, /* bad argument / shape                              */
  MRCNN_ECUDA = -2,  /* CUDA runtime or driver error, or no sm_100 device */
  MRCNN_ENCCL = -3,  /* NCCL error or libnccl not loadable                */
  MRCNN_EIO = -4,    /* anchors / weight file missing or malformed        */
  MRCNN_ESTATE = -5  /* call not valid in this state (e.g. no weights)    */
} mrcnn_status;

typedef struct mrcnn_ctx mrcnn_ctx;

/*
 * Every tunable the reference exposes, with the reference defaults.
 *   image_h/image_w        PyramidROIAlignLayer.swift:46,55-58 (imageHeight/imageWidth)
 *   architecture           README.md:87 ("resnet101" | "resnet50") -> 101 | 50
 *   num_classes            README.md:89
 *   bbox_std               ProposalLayer.swift:57,70-80; DetectionLayer.swift:55,67-77
 *   pre_nms_max_proposals  ProposalLayer.swift:59,82-84
 *   max_proposals          ProposalLayer.swift:61,85-87
 *   proposal_nms_iou       ProposalLayer.swift:63,88-90
 *   pool_size_classifier   PyramidROIAlignLayer.swift:45,51-53 (first instance)
 *   pool_size_mask         same layer, second instance (mask branch)
 *   fpn_selection_factor   PyramidROIAlignLayer.swift:98 (hard-coded 224)
 *   max_detections         DetectionLayer.swift:57,79-81
 *   detection_min_score    DetectionLayer.swift:59,82-84
 *   detection_nms_iou      DetectionLayer.swift:61,85-87
 *   mean_rgb               Conversion/task.py:73-75 (subtracted, no scale)
 *   anchors_path           MaskRCNNConfig.swift:15  (anchorsURL)
 *   main_model_path        the MaskRCNN model bundle (ViewController.swift:37)
 *   classifier_model_path  MaskRCNNConfig.swift:16  (compiledClassifierModelURL)
 *   mask_model_path        MaskRCNNConfig.swift:17  (compiledMaskModelURL)
 *   precise_masks          (no reference counterpart) numerical mode of the mask head, see the field
 * Paths may be NULL: the layer-level calls that do not need them still work
 * (proposal needs anchors; classifier/mask/predict need weights).
 */
typedef struct mrcnn_config {
  int32_t struct_size; /* sizeof(mrcnn_config), set by mrcnn_config_default */
  int32_t device;      /* CUDA device ordinal, -1 = current device          */
  int32_t image_h, image_w;
  int32_t architecture;
  int32_t num_classes;
  float bbox_std[4];
  int32_t pre_nms_max_proposals;
  int32_t max_proposals;
  float proposal_nms_iou;
  int32_t pool_size_classifier;
  int32_t pool_size_mask;
  float fpn_selection_factor;
  int32_t max_detections;
  float detection_min_score;
  float detection_nms_iou;
  float mean_rgb[3];
  int32_t max_batch; /* images per predict() call the workspace is sized for */
  int32_t precise_masks; /* 1 (default): 2-term (hi, lo) fp16 activations in the mask head: masks within 1e-4 of an fp32
                            evaluation, the tolerance of the path;  0: 1-term fp16 activations, half the mask-head tensor
                            work (~8 % more images/s), masks within 5e-4 */
  const char* anchors_path;
  const char* main_model_path;
  const char* classifier_model_path;
  const char* mask_model_path;
} mrcnn_config;

/* Fills *cfg with the reference defaults listed above (1024x1024, resnet101, 81
 * classes, 6000 -> 1000 proposals, 100 detections, ...). */
MRCNN_API void mrcnn_config_default(mrcnn_config* cfg);

MRCNN_API const char* mrcnn_version(void);

/* Lifecycle.  Replaces: custom-layer init(parameters:) (ProposalLayer.swift:65-91,
 * PyramidROIAlignLayer.swift:48-59, DetectionLayer.swift:63-88), the anchors
 * load at ProposalLayer.swift:68 and the per-call MLModel(contentsOf:) loads at
 * TimeDistributedClassifierLayer.swift:41 / TimeDistributedMaskLayer.swift:49
 * (done once here).  All device workspace is allocated here, never in *_eval. */
MRCNN_API int mrcnn_create(const mrcnn_config* cfg, mrcnn_ctx** out_ctx);
MRCNN_API void mrcnn_destroy(mrcnn_ctx* ctx);
/* Message for the last failing call on ctx (ctx may be NULL: message of the
 * last failing mrcnn_create on this thread). Never NULL. */
MRCNN_API const char* mrcnn_last_error(const mrcnn_ctx* ctx);

/* Use an externally owned CUDA stream (cudaStream_t / CUstream as void*) for
 * all work of this context; NULL restores the context's own stream.
 * Ordering contract for DEVICE pointers: the context's own stream is created
 * cudaStreamNonBlocking, i.e. it is NOT ordered with the legacy default stream
 * or with any other stream.  Work that produces an input buffer (or still reads
 * an output buffer) on another stream must have completed, or be ordered by the
 * caller (event / stream synchronisation, or by handing that stream to
 * mrcnn_set_stream), before the call.  Calls with host pointers synchronise
 * the context's stream before they return. */
MRCNN_API int mrcnn_set_stream(mrcnn_ctx* ctx, void* cuda_stream);
MRCNN_API int mrcnn_synchronize(mrcnn_ctx* ctx);

/* Anchors / weights from memory instead of files (same formats as the files:
 * anchors.bin = headerless little-endian f32 (N,4) normalised (y1,x1,y2,x2),
 * Conversion/task.py:176; weights = the blob format of DESIGN.md).
 * which: 0 = main (backbone+FPN+RPN), 1 = classifier, 2 = mask. */
MRCNN_API int mrcnn_set_anchors(mrcnn_ctx* ctx, const float* anchors, int64_t num_anchors);
MRCNN_API int mrcnn_set_weights(mrcnn_ctx* ctx, int which, const void* blob, size_t bytes);
MRCNN_API int64_t mrcnn_num_anchors(const mrcnn_ctx* ctx);
/* Anchors on demand instead of the pre-computed anchors.bin (the reference's own TODO, MaskRCNNConfig.swift:14:
 * "generate the anchors on demand based on image shape, this will save 5mb").  Host arithmetic, no device needed:
 * the Matterport rule of the reference's conversion package (scales 32..512 on strides 4..64, ratios 0.5/1/2,
 * level-major then y, x, ratio; normalised (y1,x1,y2,x2)), i.e. the content Conversion/task.py:176 writes.
 * mrcnn_anchor_count: N for an image size (261 888 at 1024x1024, 65 472 at 512x512), negative status when < 2 pixels.
 * mrcnn_generate_anchors: writes N x 4 floats; capacity = number of ROWS anchors_out can hold.  Pass the result to
 * mrcnn_set_anchors. */
MRCNN_API int64_t mrcnn_anchor_count(int image_h, int image_w);
MRCNN_API int mrcnn_generate_anchors(int image_h, int image_w, float* anchors_out, int64_t capacity);

/* ---- ProposalLayer -----------------------------------------------------
 * outputShapes (ProposalLayer.swift:97-101): (max_proposals, 4). */
MRCNN_API int mrcnn_proposal_output_shape(const mrcnn_ctx* ctx, int64_t shape_out[2]);
/* evaluate (ProposalLayer.swift:103-195).
 *   probs  [batch, N, 2] f32  (background, object) per anchor
 *   deltas [batch, N, 4] f32  (dy, dx, log dh, log dw)
 *   rois_out [batch, max_proposals, 4] f32 (y1,x1,y2,x2) normalised, zero padded
 *   keep_anchor_out (optional, may be NULL) [batch, max_proposals] i32: anchor
 *     index of every kept roi, -1 padded (parity hook for NMS keep indices)
 *   count_out (optional) [batch] i32: number of rois kept. */
MRCNN_API int mrcnn_proposal_eval(mrcnn_ctx* ctx, int batch, int64_t num_anchors,
                        const float* probs, const float* deltas,
                        float* rois_out, int32_t* keep_anchor_out,
                        int32_t* count_out);

/* ---- PyramidROIAlignLayer ------------------------------------------------
 * outputShapes (PyramidROIAlignLayer.swift:65-77): (R, C, pool, pool). */
MRCNN_API int mrcnn_pyramid_roialign_output_shape(const mrcnn_ctx* ctx, int64_t num_rois,
                                        int64_t channels, int pool,
                                        int64_t shape_out[4]);
/* evaluate (PyramidROIAlignLayer.swift:79-181, roisToInputItems :351-396,
 * copyOutput :245-274; sampling = MPSNNCropAndResizeBilinear :212-223).
 *   rois   [batch, R, roi_row_stride] f32; first 4 of each row = (y1,x1,y2,x2);
 *          roi_row_stride is 4 for proposals, 6 for detections (strides[0], :356)
 *   fmaps[l] [batch, C, H_l, W_l] f32 CHW for pyramid levels 2..5 (l = 0..3),
 *          hw[2*l] = H_l, hw[2*l+1] = W_l
 *   out    [batch, R, C, pool, pool] f32; invalid (padding) rois give zero blocks
 *   level_out (optional) [batch, R] i32: chosen level 2..5, -1 for padding. */
MRCNN_API int mrcnn_pyramid_roialign_eval(mrcnn_ctx* ctx, int batch, const float* rois,
                                int roi_row_stride, int64_t num_rois,
                                const float* const fmaps[4], const int32_t hw[8],
                                int64_t channels, int pool, float* out,
                                int32_t* level_out);

/* ---- TimeDistributedClassifierLayer ---------------------------------------
 * outputShapes (TimeDistributedClassifierLayer.swift:26-32): (R, 6).
 * evaluate (:34-91): runs the Classifier model on every pooled map then
 * argmax (first max, :177-192) + pick that class's deltas.
 *   pooled [batch, R, C, P, P] f32 (P = pool_size_classifier)
 *   out    [batch, R, 6] f32 = (dy,dx,log dh,log dw, classId, score) */
MRCNN_API int mrcnn_classifier_eval(mrcnn_ctx* ctx, int batch, int64_t num_rois,
                          const float* pooled, float* out);
/* The post-processing half alone (:50-88): probabilities [batch,R,ncls] and
 * bounding_boxes [batch,R,ncls*4] (class-major) -> out [batch,R,6]. */
MRCNN_API int mrcnn_classifier_select(mrcnn_ctx* ctx, int batch, int64_t num_rois,
                            const float* probabilities,
                            const float* bounding_boxes, float* out);

/* ---- DetectionLayer ---------------------------------------------------------
 * outputShapes (DetectionLayer.swift:94-105): (max_detections, 6). */
MRCNN_API int mrcnn_detection_output_shape(const mrcnn_ctx* ctx, int64_t shape_out[2]);
/* evaluate (DetectionLayer.swift:107-234, indicesOfRoisWithHighScores :238-276).
 *   rois [batch, R, 4] f32, classifications [batch, R, 6] f32
 *   out  [batch, max_detections, 6] f32 = (y1,x1,y2,x2,classId,score), zero padded
 *   keep_roi_out (optional) [batch, max_detections] i32: roi index per row, -1 pad
 *   count_out (optional) [batch] i32. */
MRCNN_API int mrcnn_detection_eval(mrcnn_ctx* ctx, int batch, int64_t num_rois,
                         const float* rois, const float* classifications,
                         float* out, int32_t* keep_roi_out, int32_t* count_out);

/* ---- TimeDistributedMaskLayer -------------------------------------------------
 * outputShapes (TimeDistributedMaskLayer.swift:26-37): (D, 2P, 2P).
 * evaluate (:39-91): Mask model on every non-padding pooled map, keep the plane
 * of the detection's class.
 *   pooled [batch, D, C, P, P] f32 (P = pool_size_mask), detections [batch, D, 6]
 *   out    [batch, D, 2P, 2P] f32 */
MRCNN_API int mrcnn_mask_eval(mrcnn_ctx* ctx, int batch, int64_t num_det,
                    const float* pooled, const float* detections, float* out);

/* ---- Pipeline: the `MaskRCNN` model's prediction (ViewController.swift:37-47,
 * EvaluateCommand.swift:155-171; I/O names Conversion/task.py:70-72).
 *   rgb        [batch, image_h, image_w, 3] u8 (already letter-boxed to the
 *              model size, as Vision's .scaleFit delivers it)
 *   detections [batch, max_detections, 6] f32
 *   masks      [batch, max_detections, 2*pool_size_mask, 2*pool_size_mask] f32
 * Synchronises before returning when any pointer is a host pointer. */
MRCNN_API int mrcnn_predict(mrcnn_ctx* ctx, int batch, const uint8_t* rgb,
                  float* detections, float* masks);

/* Detection.swift:23-62 + :64-99 on the device: rows with Double(score) > 0.7,
 * bounding box (x, y, w, h), and the 8-bit mask 255 - p/2*255.
 *   count_out [batch] i32; index_out/class_out [batch, D] i32;
 *   bbox_out [batch, D, 4] f64; score_out [batch, D] f64;
 *   mask_u8_out [batch, D, S*S] u8 (may be NULL). Rows >= count are zero. */
MRCNN_API int mrcnn_detections_decode(mrcnn_ctx* ctx, int batch, const float* detections,
                            const float* masks, int32_t* count_out,
                            int32_t* index_out, double* bbox_out,
                            int32_t* class_out, double* score_out,
                            uint8_t* mask_u8_out);

/* ---- Pre-processing in front of the path: Vision's `.scaleFit` (EvaluateCommand.swift:157,
 * ViewController.swift:42).  Geometry as the reference's own letter-boxing (DetectionRenderer.swift:63-75): scale to
 * fit, centred padding (black); bilinear, half-pixel centres.
 *   src [src_h, src_w, 3] u8 -> dst [image_h, image_w, 3] u8 (ready for mrcnn_predict). */
MRCNN_API int mrcnn_letterbox_eval(mrcnn_ctx* ctx, const uint8_t* src, int src_h, int src_w, uint8_t* dst);
/* out5 = {scale, new_w, new_h, pad_x, pad_y} of that mapping (host arithmetic, no GPU needed). */
MRCNN_API int mrcnn_letterbox_geometry(int src_h, int src_w, int dst_h, int dst_w, double out5[5]);
/* Inverse mapping for results: boxes (y1,x1,y2,x2) normalised to the model frame -> normalised to the source
 * image; rows of row_stride floats (4 for rois, 6 for detections), extra columns are copied.  Host arithmetic. */
MRCNN_API int mrcnn_unletterbox_boxes(int src_h, int src_w, int dst_h, int dst_w, const float* boxes,
                            int64_t n, int row_stride, float* out);

/* ---- Multi-GPU: images shard across ranks, one all-gather of the packed
 * (detections | masks) rows.  The reference is single-process
 * (EvaluateCommand.swift:166-194 loops images serially); this is the only
 * exchange step of the path.
 * nccl_unique_id: the 128-byte ncclUniqueId produced on rank 0
 * (mrcnn_nccl_unique_id) and distributed by the host launcher. */
MRCNN_API int mrcnn_nccl_unique_id(void* id_out_128);
MRCNN_API int mrcnn_comm_init(mrcnn_ctx* ctx, const void* nccl_unique_id_128, int rank,
                    int nranks);
/* predict on this rank's `batch_local` images, then all-gather:
 *   detections_all [nranks*batch_local, max_detections, 6],
 *   masks_all      [nranks*batch_local, max_detections, S, S] (device or host). */
MRCNN_API int mrcnn_predict_allgather(mrcnn_ctx* ctx, int batch_local, const uint8_t* rgb,
                            float* detections_all, float* masks_all);

/* ---- Streaming prediction: the same pipeline with two batches in flight, so that the host->device copy of batch
 * i+1 and the device->host copy of batch i-1 (own copy streams) overlap the compute of batch i and the caller
 * prepares the next batch while the GPU works.  The reference predicts image after image from a loop (EvaluateCommand.swift:166-194); this is that loop
 * with the copies taken off the critical path.
 *   mrcnn_predict_submit  enqueues one batch and returns without waiting.  Arguments as mrcnn_predict (or, with
 *                         MRCNN_SUBMIT_ALLGATHER, as mrcnn_predict_allgather).  Host buffers should be pinned and
 *                         must stay valid, and untouched, until the matching mrcnn_predict_wait returns.
 *                         MRCNN_EINVAL when two batches are already in flight.
 *   mrcnn_predict_wait    blocks until the OLDEST submitted batch is complete (its outputs are then in the buffers
 *                         that were passed to submit); results are bit-identical to mrcnn_predict.
 *   mrcnn_predict_in_flight  number of submitted, not yet waited-for batches (0..2). */
#define MRCNN_SUBMIT_ALLGATHER 1
MRCNN_API int mrcnn_predict_submit(mrcnn_ctx* ctx, int batch, const uint8_t* rgb,
                         float* detections, float* masks, int flags);
MRCNN_API int mrcnn_predict_wait(mrcnn_ctx* ctx);
MRCNN_API int mrcnn_predict_in_flight(const mrcnn_ctx* ctx);

/* ---- Instrumentation (replaces the os_signpost intervals, e.g.
 * ProposalLayer.swift:105-194).  Per-stage device milliseconds of the last
 * mrcnn_predict, names in names_out (static strings), returns count. */
MRCNN_API int mrcnn_last_stage_times(const mrcnn_ctx* ctx, int max_stages,
                           const char** names_out, float* ms_out);
/* Per-kernel-class device timing: while enabled, every kernel launch of this
 * library is bracketed by a CUDA event pair on the context's stream (consecutive launches of the same class
 * share one pair, so back-to-back kernels are timed as they run in production).
 * mrcnn_profile_read synchronises, then returns per class (static name) the summed
 * milliseconds, the number of bracketed launches and their algorithmic work
 * (bytes for the memory-bound classes, flops for conv_gemm_tcgen05) since the last
 * read, and resets the accumulators.  Returns the number of classes written. */
MRCNN_API int mrcnn_profile_enable(mrcnn_ctx* ctx, int on);
MRCNN_API int mrcnn_profile_read(mrcnn_ctx* ctx, int max_classes, const char** names_out,
                       float* ms_out, int64_t* launches_out, double* work_out);
/* Number of kernels this library launched on ctx since creation. */
MRCNN_API int64_t mrcnn_launch_count(const mrcnn_ctx* ctx);

/* ---- Internal-layout entry points used by the bench / tests to exercise the
 * fused pipeline stages in isolation (device pointers only).
 * NHWC fp16 ROIAlign: fmaps[l] [batch,H_l,W_l,C] f16, out [batch,R,P,P,C] f16. */
MRCNN_API int mrcnn_roialign_nhwc_f16(mrcnn_ctx* ctx, int batch, const float* rois,
                            int roi_row_stride, int64_t num_rois,
                            const void* const fmaps[4], const int32_t hw[8],
                            int64_t channels, int pool, void* out,
                            int32_t* level_out);

/* One NHWC fp16 convolution through the tcgen05 implicit-GEMM kernel that every
 * dense layer of the pipeline uses (test / bench hook; device pointers only).
 *   x [n,h,w,cin] f16 (cin % 64 == 0), wgt [cout,kh,kw,cin] f16, bias [cout] f32 or
 *   NULL, residual [n,h_out,w_out,ldc] f16 or NULL, out [n,h_out,w_out,ldc] f16
 *   with ldc = round_up(cout, 8). */
MRCNN_API int mrcnn_conv2d_nhwc_f16(mrcnn_ctx* ctx, const void* x, int n, int h, int w,
                          int cin, const void* wgt, const float* bias, int cout,
                          int kh, int kw, int stride, int pad,
                          const void* residual, int relu, void* out);
/* Test hook: a bottleneck block's 1x1 expansion with residual and the next block's 1x1 reduction in ONE launch
 * (csrc/conv_fused.cuh): X = relu(a * w1^T + b1 + residual) [n,h,w,n1], Y = relu(X * w2^T + b2) [n,h,w,n2];
 * a [n,h,w,c1] f16 (c1 <= 256, % 64), w1 [n1,c1] f16 (n1 % 256 == 0), w2 [n2,n1] f16 (n2 = 64 / 128 / 256).
 * Bit-identical to two mrcnn_conv2d_nhwc_f16 calls.  Device pointers only. */
MRCNN_API int mrcnn_debug_fused_expand_reduce(mrcnn_ctx* ctx, const void* a, int n, int h, int w, int c1,
                          const void* w1, const float* b1, int n1, const void* residual,
                          const void* w2, const float* b2, int n2, void* x_out, void* y_out);
/* Debug: per-CTA event trace (clock64 stamps of the producer / MMA / epilogue roles) of the next
 * mrcnn_conv2d_nhwc_f16 calls; device buffer of 148 * 3 * (2*340 + 2) u64, NULL = off (tools/trace_conv.py). */
MRCNN_API int mrcnn_debug_conv_trace(void* device_buffer);
/* Backbone + FPN + RPN only (stage-level parity hook; device pointers only):
 *   rgb [batch,H,W,3] u8 -> fmaps_out[l] [batch,H_l,W_l,256] f16 NHWC (P2..P5),
 *   probs_out [batch,N,2] f32, deltas_out [batch,N,4] f32. */
MRCNN_API int mrcnn_backbone_eval(mrcnn_ctx* ctx, int batch, const uint8_t* rgb,
                        void* const fmaps_out[4], float* probs_out,
                        float* deltas_out);

#ifdef __cplusplus
}
#endif
#endif /* MASKRCNN_CUDA_H_ */
