// maskrcnn.hpp -- C++17 host side over the C ABI of libmaskrcnn_cuda.so (include/maskrcnn_cuda.h).
//
// The reference's host code is compiled Swift; the Swift toolchain is absent in this image, so this header is
// the compiled-language mirror of the reference's plugin interface for the path (the Swift package under swift/
// is the same mirror for a machine that has `swiftc`).  Names, argument meaning and error behaviour follow
// Sources/Mask-RCNN-CoreML/ in the reference:
//
//   reference (Swift)                                            here (namespace mrcnn)
//   MLCustomLayer.init(parameters: [String: Any]) throws         Layer(const Parameters&, ContextRef)   (throws mrcnn::Error)
//   setWeightData(_:)                                            setWeightData()                          (no-op, as in every layer)
//   outputShapes(forInputShapes:) -> [[NSNumber]]                outputShapes(forInputShapes)
//   evaluate(inputs: [MLMultiArray], outputs: [MLMultiArray])    evaluate(inputs, outputs)
//   MaskRCNNConfig.defaultConfig (+ the three URL properties)    MaskRCNNConfig::defaultConfig()
//   Detection / detectionsFromFeatureValue                       Detection::detectionsFromFeatureValue
//   the generated MaskRCNN model class (ViewController.swift:37) MaskRCNN
//   Swift `throws`                                               mrcnn::Error (status + message of mrcnn_last_error)
//
// Header-only; link with -lmaskrcnn_cuda.  There is no CPU fallback: without an sm_100 device every Context
// constructor throws.  Nothing here allocates device memory per call; MultiArray is a non-owning view, like the
// MLMultiArrays Core ML hands to a custom layer (host or device pointers are both accepted by the library).
#ifndef MASKRCNN_HPP_
#define MASKRCNN_HPP_

#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "maskrcnn_cuda.h"

namespace mrcnn {

// ---- Swift `throws` -------------------------------------------------------------------------------------------
class Error : public std::runtime_error {
 public:
  Error(int status, const std::string& message)
      : std::runtime_error("[mrcnn status " + std::to_string(status) + "] " + message), status_(status) {}
  int status() const noexcept { return status_; }

 private:
  int status_;
};

// ---- [String: Any] ----------------------------------------------------------------------------------------------
// Core ML hands custom-layer parameters as Int, Double or String (Conversion/task.py:25-60 writes them).  The
// layers read them with `as? Int` / `as? Double`: a value of the wrong kind is ignored and the default stays
// (e.g. ProposalLayer.swift:82-90).  get_if below has exactly that behaviour.
using ParameterValue = std::variant<std::int64_t, double, std::string>;
using Parameters = std::map<std::string, ParameterValue>;
using Shape = std::vector<std::int64_t>;

namespace detail {
template <class T>
inline const T* get_if(const Parameters& p, const std::string& key) {
  auto it = p.find(key);
  return it == p.end() ? nullptr : std::get_if<T>(&it->second);
}
// bboxStdDev_count + bboxStdDev_<i> (ProposalLayer.swift:70-80, DetectionLayer.swift:67-77): taken only when the
// count is an Int and every item is a Double; Float(item) rounding as in the reference.
inline std::vector<float> std_dev(const Parameters& p) {
  std::vector<float> def{0.1f, 0.1f, 0.2f, 0.2f};
  const auto* n = get_if<std::int64_t>(p, "bboxStdDev_count");
  if (!n) return def;
  std::vector<float> v;
  for (std::int64_t i = 0; i < *n; ++i)
    if (const auto* d = get_if<double>(p, "bboxStdDev_" + std::to_string(i))) v.push_back(static_cast<float>(*d));
  return static_cast<std::int64_t>(v.size()) == *n ? v : def;
}
}  // namespace detail

// ---- MLMultiArray (float32) ---------------------------------------------------------------------------------------
// Dense row-major view.  `shape` is what the caller has: the reference's 5-D [seq,batch,channel,height,width] or
// the squeezed dense shape; evaluate() only uses the element counts and trailing dimensions it documents.
struct MultiArray {
  float* data = nullptr;
  Shape shape;
  MultiArray() = default;
  MultiArray(float* d, Shape s) : data(d), shape(std::move(s)) {}
  MultiArray(const float* d, Shape s) : data(const_cast<float*>(d)), shape(std::move(s)) {}
  std::int64_t count() const {
    std::int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
  // i-th dimension counted from the END with every size-1 dimension ignored (a 5-D Core ML shape and the dense
  // shape of the same array then agree)
  std::int64_t dim_from_end(int i) const {
    for (auto it = shape.rbegin(); it != shape.rend(); ++it) {
      if (*it == 1) continue;
      if (i-- == 0) return *it;
    }
    return 1;
  }
};

// ---- MaskRCNNConfig (MaskRCNNConfig.swift:10-18) --------------------------------------------------------------------
// The Swift class holds the three artefact URLs; the parameters Core ML bakes into the .mlmodel (custom-layer
// parameters, README.md:85-92 config JSON) live here too, with the Swift defaults.
class MaskRCNNConfig {
 public:
  static MaskRCNNConfig& defaultConfig() {
    static MaskRCNNConfig cfg;
    return cfg;
  }
  std::optional<std::string> anchorsURL;                  // MaskRCNNConfig.swift:15
  std::optional<std::string> compiledClassifierModelURL;  // :16
  std::optional<std::string> compiledMaskModelURL;        // :17
  std::optional<std::string> modelURL;                    // the MaskRCNN model bundle (ViewController.swift:37)
  std::string architecture = "resnet101";                 // README.md:87
  int imageHeight = 1024, imageWidth = 1024;              // README.md:88
  int numClasses = 81;                                    // README.md:89
  std::vector<float> boundingBoxRefinementStandardDeviation{0.1f, 0.1f, 0.2f, 0.2f};
  int preNMSMaxProposals = 6000;                          // ProposalLayer.swift:59
  int maxProposals = 1000;                                // ProposalLayer.swift:61
  float proposalNMSIOUThreshold = 0.7f;                   // ProposalLayer.swift:63
  int classifierPoolSize = 7, maskPoolSize = 14;          // PyramidROIAlignLayer.swift:45 (one instance each)
  int maxDetections = 100;                                // DetectionLayer.swift:57
  float scoreThreshold = 0.7f;                            // DetectionLayer.swift:59
  float detectionNMSIOUThreshold = 0.3f;                  // DetectionLayer.swift:61
  int maxBatch = 8;                                       // images per predict call the workspace is sized for
  bool preciseMasks = true;                               // mrcnn_config.precise_masks (default: masks within 1e-4 of fp32)
  int device = -1;                                        // CUDA ordinal, -1 = current

  // The C struct; the returned pointers into *this stay valid while *this is unchanged.
  mrcnn_config c_config() const {
    mrcnn_config c;
    mrcnn_config_default(&c);
    c.device = device;
    c.image_h = imageHeight;
    c.image_w = imageWidth;
    if (architecture == "resnet101") c.architecture = 101;
    else if (architecture == "resnet50") c.architecture = 50;
    else throw Error(MRCNN_EINVAL, "architecture must be resnet101 or resnet50, got " + architecture);
    c.num_classes = numClasses;
    if (boundingBoxRefinementStandardDeviation.size() != 4)
      throw Error(MRCNN_EINVAL, "boundingBoxRefinementStandardDeviation needs 4 values");
    for (int i = 0; i < 4; ++i) c.bbox_std[i] = boundingBoxRefinementStandardDeviation[static_cast<size_t>(i)];
    c.pre_nms_max_proposals = preNMSMaxProposals;
    c.max_proposals = maxProposals;
    c.proposal_nms_iou = proposalNMSIOUThreshold;
    c.pool_size_classifier = classifierPoolSize;
    c.pool_size_mask = maskPoolSize;
    c.max_detections = maxDetections;
    c.detection_min_score = scoreThreshold;
    c.detection_nms_iou = detectionNMSIOUThreshold;
    c.max_batch = maxBatch;
    c.precise_masks = preciseMasks ? 1 : 0;
    c.anchors_path = anchorsURL ? anchorsURL->c_str() : nullptr;
    c.main_model_path = modelURL ? modelURL->c_str() : nullptr;
    c.classifier_model_path = compiledClassifierModelURL ? compiledClassifierModelURL->c_str() : nullptr;
    c.mask_model_path = compiledMaskModelURL ? compiledMaskModelURL->c_str() : nullptr;
    return c;
  }
};

// ---- one mrcnn_ctx: one CUDA device + stream; not thread-safe ------------------------------------------------------
class Context {
 public:
  explicit Context(const MaskRCNNConfig& cfg = MaskRCNNConfig::defaultConfig()) : config_(cfg) {
    c_ = config_.c_config();  // pointers refer to config_ (our copy), not to the caller's object
    int st = mrcnn_create(&c_, &ctx_);
    if (st != MRCNN_OK) throw Error(st, mrcnn_last_error(nullptr));
  }
  ~Context() {
    if (ctx_) mrcnn_destroy(ctx_);
  }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;

  mrcnn_ctx* handle() const noexcept { return ctx_; }
  const mrcnn_config& c_config() const noexcept { return c_; }
  const MaskRCNNConfig& config() const noexcept { return config_; }
  void check(int status) const {
    if (status != MRCNN_OK) throw Error(status, mrcnn_last_error(ctx_));
  }
  void setAnchors(const float* anchors, std::int64_t n) { check(mrcnn_set_anchors(ctx_, anchors, n)); }
  // anchors for the configured image size, generated on demand instead of read from anchors.bin (the reference's
  // own TODO, MaskRCNNConfig.swift:14)
  void generateAnchors() {
    const std::int64_t n = mrcnn_anchor_count(c_.image_h, c_.image_w);
    if (n < 0) throw Error(static_cast<int>(n), "mrcnn_anchor_count: bad image size");
    std::vector<float> a(static_cast<size_t>(n) * 4);
    int st = mrcnn_generate_anchors(c_.image_h, c_.image_w, a.data(), n);
    if (st != MRCNN_OK) throw Error(st, "mrcnn_generate_anchors failed");
    setAnchors(a.data(), n);
  }
  // which: 0 = MaskRCNN (backbone + FPN + RPN), 1 = Classifier, 2 = Mask
  void setWeights(int which, const void* blob, size_t bytes) { check(mrcnn_set_weights(ctx_, which, blob, bytes)); }
  void setStream(void* cuda_stream) { check(mrcnn_set_stream(ctx_, cuda_stream)); }
  void synchronize() { check(mrcnn_synchronize(ctx_)); }
  std::int64_t numAnchors() const noexcept { return mrcnn_num_anchors(ctx_); }
  std::int64_t launchCount() const noexcept { return mrcnn_launch_count(ctx_); }
  // os_signpost intervals of the last predict (e.g. ProposalLayer.swift:105-194): (name, device ms)
  std::vector<std::pair<std::string, float>> stageTimes() const {
    const char* names[32];
    float ms[32];
    int n = mrcnn_last_stage_times(ctx_, 32, names, ms);
    std::vector<std::pair<std::string, float>> out;
    for (int i = 0; i < n; ++i) out.emplace_back(names[i], ms[i]);
    return out;
  }

  // The process-wide context, the analogue of every reference layer reading MaskRCNNConfig.defaultConfig
  // (ProposalLayer.swift:68).  Created on first use from the singleton's state at that moment.
  static std::shared_ptr<Context> shared() {
    static std::shared_ptr<Context> ctx = std::make_shared<Context>(MaskRCNNConfig::defaultConfig());
    return ctx;
  }

 private:
  MaskRCNNConfig config_;
  mrcnn_config c_{};
  mrcnn_ctx* ctx_ = nullptr;
};
using ContextRef = std::shared_ptr<Context>;

// ---- the five MLCustomLayer classes ----------------------------------------------------------------------------------
class CustomLayer {
 public:
  virtual ~CustomLayer() = default;
  // setWeightData(_:) is a no-op in every reference layer (e.g. ProposalLayer.swift:93-95)
  void setWeightData(const std::vector<std::vector<std::uint8_t>>& = {}) {}
  virtual std::vector<Shape> outputShapes(const std::vector<Shape>& forInputShapes) const = 0;
  virtual void evaluate(const std::vector<MultiArray>& inputs, const std::vector<MultiArray>& outputs) = 0;
  // The context the layer runs on: the one given to the constructor, else the process-wide one (created on first
  // use, so that parameter parsing and outputShapes work before any device is touched, as Core ML calls them at
  // model-load time).  The first call checks the layer's parameters against the context's configuration and
  // throws when they differ: the context holds the values the kernels run with, a layer with other parameters
  // needs its own Context.
  Context& context() {
    if (!ctx_) ctx_ = Context::shared();
    if (!validated_) {
      validate(*ctx_);
      validated_ = true;
    }
    return *ctx_;
  }

 protected:
  explicit CustomLayer(ContextRef ctx) : ctx_(std::move(ctx)) {}
  virtual void validate(const Context&) {}
  static void need(bool ok, const char* what) {
    if (!ok) throw Error(MRCNN_EINVAL, what);
  }
  static bool same_std(const mrcnn_config& c, const std::vector<float>& v) {
    if (v.size() != 4) return false;
    for (size_t i = 0; i < 4; ++i)
      if (c.bbox_std[i] != v[i]) return false;
    return true;
  }

 private:
  ContextRef ctx_;
  bool validated_ = false;
};

// ProposalLayer.swift:52-197.  inputs = {probs (N,2), deltas (N,4)} (optionally with a leading image batch),
// outputs = {rois (maxProposals,4)}.
class ProposalLayer : public CustomLayer {
 public:
  std::vector<float> boundingBoxRefinementStandardDeviation;
  int preNMSMaxProposals = 6000;
  int maxProposals = 1000;
  float nmsIOUThreshold = 0.7f;

  explicit ProposalLayer(const Parameters& parameters = {}, ContextRef ctx = nullptr) : CustomLayer(std::move(ctx)) {
    boundingBoxRefinementStandardDeviation = detail::std_dev(parameters);
    if (const auto* v = detail::get_if<std::int64_t>(parameters, "preNMSMaxProposals")) preNMSMaxProposals = static_cast<int>(*v);
    if (const auto* v = detail::get_if<std::int64_t>(parameters, "maxProposals")) maxProposals = static_cast<int>(*v);
    if (const auto* v = detail::get_if<double>(parameters, "nmsIOUThreshold")) nmsIOUThreshold = static_cast<float>(*v);
  }
  void validate(const Context& ctx) override {
    const mrcnn_config& c = ctx.c_config();
    need(c.pre_nms_max_proposals == preNMSMaxProposals && c.max_proposals == maxProposals &&
             c.proposal_nms_iou == nmsIOUThreshold && same_std(c, boundingBoxRefinementStandardDeviation),
         "ProposalLayer parameters differ from the context configuration");
  }
  // ProposalLayer.swift:97-101: the deltas' shape with dimension 0 replaced by maxProposals
  std::vector<Shape> outputShapes(const std::vector<Shape>& forInputShapes) const override {
    need(forInputShapes.size() >= 2 && !forInputShapes[1].empty(), "ProposalLayer.outputShapes needs 2 input shapes");
    Shape out = forInputShapes[1];
    out[0] = maxProposals;
    return {out};
  }
  void evaluate(const std::vector<MultiArray>& inputs, const std::vector<MultiArray>& outputs) override {
    evaluate(inputs, outputs, nullptr, nullptr);
  }
  // keepAnchor [batch, maxProposals] / count [batch]: optional parity hooks (anchor index of every kept roi)
  void evaluate(const std::vector<MultiArray>& inputs, const std::vector<MultiArray>& outputs, std::int32_t* keepAnchor,
                std::int32_t* count) {
    need(inputs.size() == 2 && outputs.size() == 1, "ProposalLayer.evaluate takes 2 inputs and 1 output");
    Context& c = context();
    const MultiArray& probs = inputs[0];
    const MultiArray& deltas = inputs[1];
    const std::int64_t n = c.numAnchors();
    need(n > 0, "ProposalLayer needs anchors (MaskRCNNConfig.anchorsURL or Context::setAnchors)");
    need(probs.count() % (2 * n) == 0 && probs.count() > 0, "probs must hold (batch, N, 2) values");
    const std::int64_t batch = probs.count() / (2 * n);
    need(deltas.count() == batch * n * 4, "deltas must hold (batch, N, 4) values");
    need(outputs[0].count() == batch * maxProposals * 4, "rois output must hold (batch, maxProposals, 4) values");
    c.check(mrcnn_proposal_eval(c.handle(), static_cast<int>(batch), n, probs.data, deltas.data, outputs[0].data,
                                    keepAnchor, count));
  }
};

// PyramidROIAlignLayer.swift:40-183.  inputs = {rois (R,4|6), P2, P3, P4, P5 as (C,H,W)}, outputs = {(R,C,pool,pool)}.
class PyramidROIAlignLayer : public CustomLayer {
 public:
  int poolSize = 7;
  int imageWidth = 1024, imageHeight = 1024;

  explicit PyramidROIAlignLayer(const Parameters& parameters = {}, ContextRef ctx = nullptr) : CustomLayer(std::move(ctx)) {
    if (const auto* v = detail::get_if<std::int64_t>(parameters, "poolSize")) poolSize = static_cast<int>(*v);
    // the reference reads imageWidth/imageHeight `as? CGFloat` although the converter writes
    // Ints, so its 1024 default always wins; intended behaviour = the configured model size.  Both kinds are
    // accepted here and must agree with the context (checked at the first evaluate).
    if (const auto* v = detail::get_if<std::int64_t>(parameters, "imageWidth")) imageWidth = static_cast<int>(*v);
    if (const auto* v = detail::get_if<double>(parameters, "imageWidth")) imageWidth = static_cast<int>(*v);
    if (const auto* v = detail::get_if<std::int64_t>(parameters, "imageHeight")) imageHeight = static_cast<int>(*v);
    if (const auto* v = detail::get_if<double>(parameters, "imageHeight")) imageHeight = static_cast<int>(*v);
    explicitSize_ = parameters.count("imageWidth") != 0 || parameters.count("imageHeight") != 0;
  }
  void validate(const Context& ctx) override {
    const mrcnn_config& c = ctx.c_config();
    if (!explicitSize_) {
      imageWidth = c.image_w;
      imageHeight = c.image_h;
    }
    need(c.image_w == imageWidth && c.image_h == imageHeight, "PyramidROIAlignLayer image size differs from the context configuration");
    need(poolSize >= 1, "poolSize must be >= 1");
  }
  // PyramidROIAlignLayer.swift:65-77
  std::vector<Shape> outputShapes(const std::vector<Shape>& forInputShapes) const override {
    need(forInputShapes.size() >= 2 && forInputShapes[0].size() >= 2 && forInputShapes[1].size() >= 3,
         "PyramidROIAlignLayer.outputShapes needs the rois and one feature-map shape (5-D)");
    const Shape& rois = forInputShapes[0];
    const Shape& fmap = forInputShapes[1];
    return {{rois[0], rois[1], fmap[2], poolSize, poolSize}};
  }
  void evaluate(const std::vector<MultiArray>& inputs, const std::vector<MultiArray>& outputs) override {
    evaluate(inputs, outputs, 1, nullptr);
  }
  // batch: number of images stacked in every array; level [batch, R]: optional parity hook (chosen level 2..5)
  void evaluate(const std::vector<MultiArray>& inputs, const std::vector<MultiArray>& outputs, int batch, std::int32_t* level) {
    need(inputs.size() == 5 && outputs.size() == 1, "PyramidROIAlignLayer.evaluate takes rois + 4 feature maps and 1 output");
    need(batch >= 1, "batch must be >= 1");
    Context& c = context();
    const MultiArray& rois = inputs[0];
    // row stride = strides[0] of the rois array (PyramidROIAlignLayer.swift:356): 4 for proposals, 6 for detections
    const std::int64_t stride = rois.dim_from_end(0);
    need(stride == 4 || stride == 6, "rois rows must have 4 or 6 values");
    const std::int64_t r = rois.count() / (stride * batch);
    const float* maps[4];
    std::int32_t hw[8];
    std::int64_t channels = 0;
    for (int l = 0; l < 4; ++l) {
      const MultiArray& f = inputs[static_cast<size_t>(l) + 1];
      const std::int64_t w = f.shape.empty() ? 0 : f.shape[f.shape.size() - 1];
      const std::int64_t h = f.shape.size() < 2 ? 0 : f.shape[f.shape.size() - 2];
      need(h > 0 && w > 0 && f.count() % (h * w * batch) == 0, "feature maps must be (batch, C, H, W)");
      const std::int64_t ch = f.count() / (h * w * batch);
      need(l == 0 || ch == channels, "feature maps must have the same channel count");
      channels = ch;
      maps[l] = f.data;
      hw[2 * l] = static_cast<std::int32_t>(h);
      hw[2 * l + 1] = static_cast<std::int32_t>(w);
    }
    need(outputs[0].count() == batch * r * channels * poolSize * poolSize, "output must hold (batch, R, C, pool, pool) values");
    c.check(mrcnn_pyramid_roialign_eval(c.handle(), batch, rois.data, static_cast<int>(stride), r, maps, hw, channels,
                                            poolSize, outputs[0].data, level));
  }

 private:
  bool explicitSize_ = false;
};

// TimeDistributedClassifierLayer.swift:14-92.  inputs = {pooled (R,C,P,P)}, outputs = {(R,6)}.
class TimeDistributedClassifierLayer : public CustomLayer {
 public:
  explicit TimeDistributedClassifierLayer(const Parameters& = {}, ContextRef ctx = nullptr) : CustomLayer(std::move(ctx)) {}
  // :26-32
  std::vector<Shape> outputShapes(const std::vector<Shape>& forInputShapes) const override {
    need(!forInputShapes.empty() && forInputShapes[0].size() >= 2, "TimeDistributedClassifierLayer.outputShapes needs a 5-D shape");
    const Shape& s = forInputShapes[0];
    return {{s[0], s[1], 1, 1, 6}};
  }
  void evaluate(const std::vector<MultiArray>& inputs, const std::vector<MultiArray>& outputs) override {
    need(inputs.size() == 1 && outputs.size() == 1, "TimeDistributedClassifierLayer.evaluate takes 1 input and 1 output");
    Context& c = context();
    const std::int64_t rows = outputs[0].count() / 6;  // batch * R
    need(rows > 0 && outputs[0].count() == rows * 6 && inputs[0].count() % rows == 0, "output must hold (R, 6) values");
    // the library takes (batch, R); rows of different images are independent, so batch = 1, R = rows is the same call
    c.check(mrcnn_classifier_eval(c.handle(), 1, rows, inputs[0].data, outputs[0].data));
  }
  // the post-processing half alone (:50-88) on explicit Classifier outputs: probabilities (R,ncls), bounding_boxes (R,ncls*4)
  void select(const MultiArray& probabilities, const MultiArray& boundingBoxes, const MultiArray& out) {
    Context& c = context();
    const std::int64_t ncls = c.c_config().num_classes;
    need(probabilities.count() % ncls == 0, "probabilities must hold (R, numClasses) values");
    const std::int64_t rows = probabilities.count() / ncls;
    need(boundingBoxes.count() == rows * ncls * 4 && out.count() == rows * 6, "bounding_boxes (R, 4*numClasses), out (R, 6)");
    c.check(mrcnn_classifier_select(c.handle(), 1, rows, probabilities.data, boundingBoxes.data, out.data));
  }
};

// DetectionLayer.swift:52-236.  inputs = {rois (R,4), classifications (R,6)}, outputs = {(maxDetections,6)}.
class DetectionLayer : public CustomLayer {
 public:
  std::vector<float> boundingBoxRefinementStandardDeviation;
  int maxDetections = 100;
  float lowConfidenceScoreThreshold = 0.7f;
  float nmsIOUThreshold = 0.3f;

  explicit DetectionLayer(const Parameters& parameters = {}, ContextRef ctx = nullptr) : CustomLayer(std::move(ctx)) {
    boundingBoxRefinementStandardDeviation = detail::std_dev(parameters);
    if (const auto* v = detail::get_if<std::int64_t>(parameters, "maxDetections")) maxDetections = static_cast<int>(*v);
    if (const auto* v = detail::get_if<double>(parameters, "scoreThreshold")) lowConfidenceScoreThreshold = static_cast<float>(*v);
    if (const auto* v = detail::get_if<double>(parameters, "nmsIOUThreshold")) nmsIOUThreshold = static_cast<float>(*v);
  }
  void validate(const Context& ctx) override {
    const mrcnn_config& c = ctx.c_config();
    need(c.max_detections == maxDetections && c.detection_min_score == lowConfidenceScoreThreshold &&
             c.detection_nms_iou == nmsIOUThreshold && same_std(c, boundingBoxRefinementStandardDeviation),
         "DetectionLayer parameters differ from the context configuration");
  }
  // DetectionLayer.swift:94-105
  std::vector<Shape> outputShapes(const std::vector<Shape>& forInputShapes) const override {
    need(!forInputShapes.empty() && forInputShapes[0].size() >= 2, "DetectionLayer.outputShapes needs the rois shape (5-D)");
    return {{maxDetections, forInputShapes[0][1], 6, 1, 1}};
  }
  void evaluate(const std::vector<MultiArray>& inputs, const std::vector<MultiArray>& outputs) override {
    evaluate(inputs, outputs, nullptr, nullptr);
  }
  // keepRoi [batch, maxDetections] / count [batch]: optional parity hooks (roi index of every output row)
  void evaluate(const std::vector<MultiArray>& inputs, const std::vector<MultiArray>& outputs, std::int32_t* keepRoi,
                std::int32_t* count) {
    need(inputs.size() == 2 && outputs.size() == 1, "DetectionLayer.evaluate takes 2 inputs and 1 output");
    Context& c = context();
    const std::int64_t per_image = static_cast<std::int64_t>(maxDetections) * 6;
    need(outputs[0].count() > 0 && outputs[0].count() % per_image == 0, "output must hold (batch, maxDetections, 6) values");
    const std::int64_t batch = outputs[0].count() / per_image;
    need(inputs[0].count() % (4 * batch) == 0, "rois must hold (batch, R, 4) values");
    const std::int64_t r = inputs[0].count() / (4 * batch);
    need(inputs[1].count() == batch * r * 6, "classifications must hold (batch, R, 6) values");
    c.check(mrcnn_detection_eval(c.handle(), static_cast<int>(batch), r, inputs[0].data, inputs[1].data,
                                     outputs[0].data, keepRoi, count));
  }
};

// TimeDistributedMaskLayer.swift:14-92.  inputs = {pooled (D,C,P,P), detections (D,6)}, outputs = {(D,2P,2P)}.
class TimeDistributedMaskLayer : public CustomLayer {
 public:
  explicit TimeDistributedMaskLayer(const Parameters& = {}, ContextRef ctx = nullptr) : CustomLayer(std::move(ctx)) {}
  // :26-37
  std::vector<Shape> outputShapes(const std::vector<Shape>& forInputShapes) const override {
    need(!forInputShapes.empty() && forInputShapes[0].size() == 5, "TimeDistributedMaskLayer.outputShapes needs a 5-D shape");
    const Shape& s = forInputShapes[0];
    return {{1, s[1], s[0], s[3] * 2, s[4] * 2}};
  }
  void evaluate(const std::vector<MultiArray>& inputs, const std::vector<MultiArray>& outputs) override {
    need(inputs.size() == 2 && outputs.size() == 1, "TimeDistributedMaskLayer.evaluate takes 2 inputs and 1 output");
    Context& c = context();
    const std::int64_t per_image = static_cast<std::int64_t>(c.c_config().max_detections) * 6;
    need(inputs[1].count() > 0 && inputs[1].count() % per_image == 0, "detections must hold (batch, maxDetections, 6) values");
    const std::int64_t batch = inputs[1].count() / per_image;
    const std::int64_t d = c.c_config().max_detections;
    const std::int64_t s = 2 * static_cast<std::int64_t>(c.c_config().pool_size_mask);
    need(outputs[0].count() == batch * d * s * s, "output must hold (batch, D, 2P, 2P) values");
    c.check(mrcnn_mask_eval(c.handle(), static_cast<int>(batch), d, inputs[0].data, inputs[1].data, outputs[0].data));
  }
};

// ---- Detection (Detection.swift:15-99) ------------------------------------------------------------------------------------
struct Rect {  // CGRect: origin + size, normalised to the model frame
  double x = 0, y = 0, width = 0, height = 0;
};

struct Detection {
  int index = 0;
  Rect boundingBox;
  int classId = 0;
  double score = 0;
  int maskSize = 0;                // side of the square mask (2 * maskPoolSize), 0 when no mask was passed
  std::vector<std::uint8_t> mask;  // 8-bit, 255 - p/2*255 (Detection.swift:83-85)

  // Detection.swift:23-62 (+ :64-99), evaluated on the device for one image: rows with Double(score) > 0.7.
  // detections (maxDetections,6); masks (maxDetections,S,S) or nullptr.
  static std::vector<Detection> detectionsFromFeatureValue(const MultiArray& featureValue, const MultiArray* maskFeatureValue,
                                                          ContextRef ctx = nullptr) {
    if (!ctx) ctx = Context::shared();
    const int d = ctx->c_config().max_detections;
    const int s = 2 * ctx->c_config().pool_size_mask;
    if (featureValue.data == nullptr) return {};  // `guard let rawDetections ... else { return [] }`
    if (featureValue.count() != static_cast<std::int64_t>(d) * 6) throw Error(MRCNN_EINVAL, "detections must hold (maxDetections, 6) values");
    if (maskFeatureValue && maskFeatureValue->count() != static_cast<std::int64_t>(d) * s * s)
      throw Error(MRCNN_EINVAL, "mask must hold (maxDetections, S, S) values");
    std::int32_t count = 0;
    std::vector<std::int32_t> index(static_cast<size_t>(d)), cls(static_cast<size_t>(d));
    std::vector<double> bbox(static_cast<size_t>(d) * 4), score(static_cast<size_t>(d));
    std::vector<std::uint8_t> mu8(maskFeatureValue ? static_cast<size_t>(d) * s * s : 0);
    ctx->check(mrcnn_detections_decode(ctx->handle(), 1, featureValue.data, maskFeatureValue ? maskFeatureValue->data : nullptr,
                                       &count, index.data(), bbox.data(), cls.data(), score.data(),
                                       maskFeatureValue ? mu8.data() : nullptr));
    std::vector<Detection> out(static_cast<size_t>(count));
    for (size_t i = 0; i < out.size(); ++i) {
      Detection& o = out[i];
      o.index = index[i];
      o.boundingBox = Rect{bbox[4 * i], bbox[4 * i + 1], bbox[4 * i + 2], bbox[4 * i + 3]};
      o.classId = cls[i];
      o.score = score[i];
      if (maskFeatureValue) {
        o.maskSize = s;
        o.mask.assign(mu8.begin() + static_cast<std::ptrdiff_t>(i * s * s), mu8.begin() + static_cast<std::ptrdiff_t>((i + 1) * s * s));
      }
    }
    return out;
  }
};

// ---- the model class (ViewController.swift:37-47; I/O names Conversion/task.py:70-72) -------------------------------------
class MaskRCNN {
 public:
  struct Output {  // MaskRCNNOutput: "detections" (D,6) and "mask" (D,S,S), per image
    std::vector<float> detections;
    std::vector<float> mask;
  };

  explicit MaskRCNN(const MaskRCNNConfig& configuration = MaskRCNNConfig::defaultConfig())
      : ctx_(std::make_shared<Context>(configuration)) {
    if (!configuration.anchorsURL) ctx_->generateAnchors();  // no anchors.bin configured: generated for the input size
  }
  explicit MaskRCNN(ContextRef ctx) : ctx_(std::move(ctx)) {
    if (!ctx_) throw Error(MRCNN_EINVAL, "MaskRCNN needs a context");
  }

  const ContextRef& context() const noexcept { return ctx_; }
  int maxDetections() const noexcept { return ctx_->c_config().max_detections; }
  int maskSize() const noexcept { return 2 * ctx_->c_config().pool_size_mask; }
  size_t imageBytes() const noexcept { return static_cast<size_t>(ctx_->c_config().image_h) * ctx_->c_config().image_w * 3; }
  size_t detectionFloats() const noexcept { return static_cast<size_t>(maxDetections()) * 6; }
  size_t maskFloats() const noexcept { return static_cast<size_t>(maxDetections()) * maskSize() * maskSize(); }

  // images [batch, H, W, 3] u8 (letter-boxed to the model size) -> detections [batch, D, 6], masks [batch, D, S, S];
  // host or device pointers (host outputs imply a stream synchronisation before returning)
  void predictionBatch(int batch, const std::uint8_t* images, float* detections, float* masks) {
    ctx_->check(mrcnn_predict(ctx_->handle(), batch, images, detections, masks));
  }
  // one image, host buffers
  Output prediction(const std::uint8_t* image) {
    Output o;
    o.detections.resize(detectionFloats());
    o.mask.resize(maskFloats());
    predictionBatch(1, image, o.detections.data(), o.mask.data());
    return o;
  }
  // image -> [Detection] with score > 0.7, the call pattern of ViewController.swift:163-187
  std::vector<Detection> predict(const std::uint8_t* image) {
    Output o = prediction(image);
    MultiArray det(o.detections.data(), {maxDetections(), 6});
    MultiArray msk(o.mask.data(), {maxDetections(), maskSize(), maskSize()});
    return Detection::detectionsFromFeatureValue(det, &msk, ctx_);
  }

  // streaming: the per-image loop of EvaluateCommand.swift:166-194 with two batches in flight
  void submit(int batch, const std::uint8_t* images, float* detections, float* masks, bool allgather = false) {
    ctx_->check(mrcnn_predict_submit(ctx_->handle(), batch, images, detections, masks, allgather ? MRCNN_SUBMIT_ALLGATHER : 0));
  }
  void wait() { ctx_->check(mrcnn_predict_wait(ctx_->handle())); }
  int inFlight() const noexcept { return mrcnn_predict_in_flight(ctx_->handle()); }

  // multi-GPU (one process per GPU): rank 0 makes the id, the launcher distributes it
  static std::vector<std::uint8_t> ncclUniqueId() {
    std::vector<std::uint8_t> id(128);
    int st = mrcnn_nccl_unique_id(id.data());
    if (st != MRCNN_OK) throw Error(st, mrcnn_last_error(nullptr));
    return id;
  }
  void commInit(const std::vector<std::uint8_t>& id, int rank, int nranks) {
    if (id.size() != 128) throw Error(MRCNN_EINVAL, "ncclUniqueId must be 128 bytes");
    ctx_->check(mrcnn_comm_init(ctx_->handle(), id.data(), rank, nranks));
  }
  void predictionAllGather(int batchLocal, const std::uint8_t* images, float* detectionsAll, float* masksAll) {
    ctx_->check(mrcnn_predict_allgather(ctx_->handle(), batchLocal, images, detectionsAll, masksAll));
  }

  // Vision's .scaleFit in front of the model (EvaluateCommand.swift:157) and its inverse for boxes
  void letterbox(const std::uint8_t* src, int srcH, int srcW, std::uint8_t* dst) {
    ctx_->check(mrcnn_letterbox_eval(ctx_->handle(), src, srcH, srcW, dst));
  }
  void unletterbox(int srcH, int srcW, const float* rows, std::int64_t n, int rowStride, float* out) const {
    ctx_->check(mrcnn_unletterbox_boxes(srcH, srcW, ctx_->c_config().image_h, ctx_->c_config().image_w, rows, n, rowStride, out));
  }

 private:
  ContextRef ctx_;
};

}  // namespace mrcnn
#endif  // MASKRCNN_HPP_
