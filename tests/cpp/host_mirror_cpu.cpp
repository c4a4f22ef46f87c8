// CPU checks of include/maskrcnn.hpp (no device needed): parameter parsing with the reference's `as? Int` /
// `as? Double` behaviour, outputShapes of the five custom layers, MaskRCNNConfig -> mrcnn_config, and the loud
// failure when there is no sm_100 device.  Built and run by tests/test_cpp_host_cpu.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "maskrcnn.hpp"

using namespace mrcnn;

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                        \
    }                                                                      \
  } while (0)

int main(int argc, char** argv) {
  const bool expect_gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;

  // ---- ProposalLayer.init(parameters:) (ProposalLayer.swift:65-91)
  {
    ProposalLayer d;
    CHECK(d.preNMSMaxProposals == 6000 && d.maxProposals == 1000 && d.nmsIOUThreshold == 0.7f);
    CHECK((d.boundingBoxRefinementStandardDeviation == std::vector<float>{0.1f, 0.1f, 0.2f, 0.2f}));
    Parameters p{{"bboxStdDev_count", std::int64_t(4)}, {"bboxStdDev_0", 0.5},      {"bboxStdDev_1", 0.25},
                 {"bboxStdDev_2", 0.125},               {"bboxStdDev_3", 0.0625},   {"preNMSMaxProposals", std::int64_t(600)},
                 {"maxProposals", std::int64_t(100)},   {"nmsIOUThreshold", 0.5}};
    ProposalLayer l(p);
    CHECK(l.preNMSMaxProposals == 600 && l.maxProposals == 100 && l.nmsIOUThreshold == 0.5f);
    CHECK((l.boundingBoxRefinementStandardDeviation == std::vector<float>{0.5f, 0.25f, 0.125f, 0.0625f}));
    // wrong kinds are ignored (`as? Int` on a Double, `as? Double` on an Int): defaults stay
    Parameters w{{"preNMSMaxProposals", 600.0}, {"maxProposals", std::string("100")}, {"nmsIOUThreshold", std::int64_t(1)}};
    ProposalLayer k(w);
    CHECK(k.preNMSMaxProposals == 6000 && k.maxProposals == 1000 && k.nmsIOUThreshold == 0.7f);
    // an incomplete std-dev list keeps the default (:77-79)
    Parameters q{{"bboxStdDev_count", std::int64_t(4)}, {"bboxStdDev_0", 0.5}};
    CHECK((ProposalLayer(q).boundingBoxRefinementStandardDeviation == std::vector<float>{0.1f, 0.1f, 0.2f, 0.2f}));
    // outputShapes (:97-101): inputShapes[1] with dim 0 := maxProposals
    auto s = l.outputShapes({{261888, 1, 2, 1, 1}, {261888, 1, 4, 1, 1}});
    CHECK(s.size() == 1 && (s[0] == Shape{100, 1, 4, 1, 1}));
    l.setWeightData();
  }
  // ---- PyramidROIAlignLayer (PyramidROIAlignLayer.swift:48-77)
  {
    PyramidROIAlignLayer l(Parameters{{"poolSize", std::int64_t(14)}});
    CHECK(l.poolSize == 14);
    CHECK(PyramidROIAlignLayer().poolSize == 7);
    auto s = l.outputShapes({{100, 1, 6, 1, 1}, {1, 1, 256, 256, 256}});
    CHECK((s[0] == Shape{100, 1, 256, 14, 14}));
  }
  // ---- TimeDistributedClassifierLayer (:26-32), TimeDistributedMaskLayer (:26-37)
  {
    auto s = TimeDistributedClassifierLayer().outputShapes({{1000, 1, 256, 7, 7}});
    CHECK((s[0] == Shape{1000, 1, 1, 1, 6}));
    auto m = TimeDistributedMaskLayer().outputShapes({{100, 1, 256, 14, 14}, {100, 1, 6, 1, 1}});
    CHECK((m[0] == Shape{1, 1, 100, 28, 28}));
  }
  // ---- DetectionLayer (DetectionLayer.swift:63-105)
  {
    DetectionLayer d;
    CHECK(d.maxDetections == 100 && d.lowConfidenceScoreThreshold == 0.7f && d.nmsIOUThreshold == 0.3f);
    DetectionLayer l(Parameters{{"maxDetections", std::int64_t(50)}, {"scoreThreshold", 0.5}, {"nmsIOUThreshold", 0.25}});
    CHECK(l.maxDetections == 50 && l.lowConfidenceScoreThreshold == 0.5f && l.nmsIOUThreshold == 0.25f);
    auto s = l.outputShapes({{1000, 1, 4, 1, 1}, {1000, 1, 1, 1, 6}});
    CHECK((s[0] == Shape{50, 1, 6, 1, 1}));
  }
  // ---- MultiArray
  {
    float buf[24] = {0};
    MultiArray a(buf, {4, 1, 6, 1, 1}), b(buf, {4, 6}), c5(buf, {1, 1, 2, 3, 4});
    CHECK(a.count() == 24 && b.count() == 24 && a.dim_from_end(0) == 6 && b.dim_from_end(0) == 6 && a.dim_from_end(1) == 4);
    CHECK(c5.dim_from_end(0) == 4 && c5.dim_from_end(1) == 3 && c5.dim_from_end(2) == 2 && c5.dim_from_end(3) == 1);
  }
  // ---- MaskRCNNConfig -> mrcnn_config
  {
    MaskRCNNConfig& g = MaskRCNNConfig::defaultConfig();
    CHECK(&g == &MaskRCNNConfig::defaultConfig());
    CHECK(!g.anchorsURL && !g.compiledClassifierModelURL && !g.compiledMaskModelURL);   // `URL?` = nil
    mrcnn_config ref;
    mrcnn_config_default(&ref);
    mrcnn_config c = g.c_config();
    CHECK(c.struct_size == (std::int32_t)sizeof(mrcnn_config) && c.image_h == 1024 && c.image_w == 1024 && c.architecture == 101);
    CHECK(c.num_classes == ref.num_classes && c.pre_nms_max_proposals == ref.pre_nms_max_proposals && c.max_proposals == ref.max_proposals);
    CHECK(c.proposal_nms_iou == ref.proposal_nms_iou && c.detection_min_score == ref.detection_min_score && c.detection_nms_iou == ref.detection_nms_iou);
    CHECK(c.pool_size_classifier == ref.pool_size_classifier && c.pool_size_mask == ref.pool_size_mask && c.max_detections == ref.max_detections);
    CHECK(c.fpn_selection_factor == ref.fpn_selection_factor && c.max_batch == ref.max_batch && c.precise_masks == 1);
    for (int i = 0; i < 4; ++i) CHECK(c.bbox_std[i] == ref.bbox_std[i]);
    for (int i = 0; i < 3; ++i) CHECK(c.mean_rgb[i] == ref.mean_rgb[i]);
    CHECK(!c.anchors_path && !c.main_model_path && !c.classifier_model_path && !c.mask_model_path);
    MaskRCNNConfig s;
    s.architecture = "resnet50";
    s.imageHeight = s.imageWidth = 512;
    s.maxProposals = 300;
    s.anchorsURL = "products/anchors.bin";
    mrcnn_config cs = s.c_config();
    CHECK(cs.architecture == 50 && cs.image_h == 512 && cs.max_proposals == 300 && std::strcmp(cs.anchors_path, "products/anchors.bin") == 0);
    s.architecture = "vgg16";
    bool threw = false;
    try { (void)s.c_config(); } catch (const Error& e) { threw = e.status() == MRCNN_EINVAL; }
    CHECK(threw);
  }
  // ---- no CPU fallback: without a device the first touch of a context throws with the library's message
  if (!expect_gpu) {
    bool threw = false;
    try {
      MaskRCNN model;
    } catch (const Error& e) {
      threw = e.status() == MRCNN_ECUDA && std::strstr(e.what(), "mrcnn_create") != nullptr;
      std::printf("no device: %s\n", e.what());
    }
    CHECK(threw);
    threw = false;
    float rois[8] = {0}, cls[12] = {0}, out[600] = {0};
    try {
      DetectionLayer l;
      l.evaluate({MultiArray(rois, {2, 4}), MultiArray(cls, {2, 6})}, {MultiArray(out, {100, 6})});
    } catch (const Error& e) {
      threw = e.status() == MRCNN_ECUDA;
    }
    CHECK(threw);
    // a missing anchors file is the reference's `try Data(contentsOf:)` failure (ProposalLayer.swift:68)
  }
  std::printf("host mirror CPU checks ok\n");
  return 0;
}
