// C++ caller through include/maskrcnn.hpp, shaped like the reference's own call sites: configure the
// MaskRCNNConfig singleton with the artefact locations (EvaluateCommand.swift:61-64), build the model once
// (ViewController.swift:37), then predict image after image (EvaluateCommand.swift:166-194) -- here as a stream of
// batches with two in flight -- and turn the last result into [Detection] (ViewController.swift:163-187).
//
//   g++ -std=c++17 -Iinclude examples/predict.cpp -Lmask-rcnn-coreml_b200 -lmaskrcnn_cuda -o predict_cpp
//   LD_LIBRARY_PATH=mask-rcnn-coreml_b200 ./predict_cpp products/ 8 4
// products/ holds anchors.bin, MaskRCNN.mrcnnw, Classifier.mrcnnw, Mask.mrcnnw (weights.write_products or
// tools/import_mlmodel.py).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "maskrcnn.hpp"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <products dir> [batch] [n_batches]\n", argv[0]);
    return 2;
  }
  const std::string dir = argv[1];
  const int batch = argc > 2 ? std::atoi(argv[2]) : 8, n_batches = argc > 3 ? std::atoi(argv[3]) : 4;
  try {
    mrcnn::MaskRCNNConfig& cfg = mrcnn::MaskRCNNConfig::defaultConfig();
    cfg.anchorsURL = dir + "/anchors.bin";
    cfg.modelURL = dir + "/MaskRCNN.mrcnnw";
    cfg.compiledClassifierModelURL = dir + "/Classifier.mrcnnw";
    cfg.compiledMaskModelURL = dir + "/Mask.mrcnnw";
    cfg.maxBatch = batch;
    mrcnn::MaskRCNN model;  // reads MaskRCNNConfig.defaultConfig, loads anchors + the three weight files once
    std::printf("%s\n", mrcnn_version());

    const size_t b = static_cast<size_t>(batch);
    std::vector<std::uint8_t> img[2];
    std::vector<float> det[2], msk[2];
    for (int k = 0; k < 2; ++k) {  // two batches in flight: two sets of host buffers
      img[k].resize(b * model.imageBytes());
      det[k].resize(b * model.detectionFloats());
      msk[k].resize(b * model.maskFloats());
      for (size_t i = 0; i < img[k].size(); ++i) img[k][i] = static_cast<std::uint8_t>((i * 2654435761u + static_cast<unsigned>(k)) >> 24);
    }
    for (int i = 0; i < n_batches; ++i) {
      model.submit(batch, img[i & 1].data(), det[i & 1].data(), msk[i & 1].data());
      if (i >= 1) model.wait();
    }
    model.wait();

    const int last = (n_batches - 1) & 1;
    for (int im = 0; im < batch; ++im) {
      mrcnn::MultiArray d(det[last].data() + static_cast<size_t>(im) * model.detectionFloats(), {model.maxDetections(), 6});
      mrcnn::MultiArray m(msk[last].data() + static_cast<size_t>(im) * model.maskFloats(),
                          {model.maxDetections(), model.maskSize(), model.maskSize()});
      auto detections = mrcnn::Detection::detectionsFromFeatureValue(d, &m, model.context());
      std::printf("image %d: %zu detections\n", im, detections.size());
      for (const auto& x : detections)
        std::printf("  #%d class %d score %.4f box (x %.4f, y %.4f, w %.4f, h %.4f)\n", x.index, x.classId, x.score, x.boundingBox.x,
                    x.boundingBox.y, x.boundingBox.width, x.boundingBox.height);
    }
    for (const auto& s : model.context()->stageTimes()) std::printf("  %-40s %.3f ms\n", s.first.c_str(), s.second);
  } catch (const mrcnn::Error& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
