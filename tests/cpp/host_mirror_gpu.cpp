// GPU parity of the C++ host mirror (include/maskrcnn.hpp): the five custom layers + Detection decoding are driven
// through the reference-shaped classes on the inputs of tests/golden/*.npz (dumped as raw little-endian files by
// tests/test_zz_cpp_host_gpu.py, which compares the outputs written here bit for bit with the golden arrays).
//   host_mirror_gpu <dir>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "maskrcnn.hpp"

using namespace mrcnn;

static std::string g_dir;

template <class T>
static std::vector<T> load(const std::string& name, size_t expect) {
  std::ifstream f(g_dir + "/" + name, std::ios::binary);
  std::vector<T> v(expect);
  if (!f || !f.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(expect * sizeof(T))) || f.peek() != EOF) {
    std::fprintf(stderr, "cannot read %zu values from %s\n", expect, name.c_str());
    std::exit(2);
  }
  return v;
}
template <class T>
static void save(const std::string& name, const std::vector<T>& v) {
  std::ofstream f(g_dir + "/" + name, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
  if (!f) {
    std::fprintf(stderr, "cannot write %s\n", name.c_str());
    std::exit(2);
  }
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  g_dir = argv[1];
  try {
    // ---- ProposalLayer: 128x128 model, 4092 anchors, 600 -> 100 (tests/golden/make_golden.py)
    {
      MaskRCNNConfig cfg;
      cfg.imageHeight = cfg.imageWidth = 128;
      cfg.preNMSMaxProposals = 600;
      cfg.maxProposals = 100;
      cfg.maxBatch = 1;
      auto ctx = std::make_shared<Context>(cfg);
      const std::int64_t n = 4092;
      auto anchors = load<float>("anchors.f32", n * 4);
      ctx->setAnchors(anchors.data(), n);
      ProposalLayer layer(Parameters{{"preNMSMaxProposals", std::int64_t(600)}, {"maxProposals", std::int64_t(100)}}, ctx);
      auto deltas = load<float>("deltas.f32", n * 4);
      for (const char* tag : {"", "tie_"}) {
        auto probs = load<float>(std::string(tag) + "probs.f32", n * 2);
        std::vector<float> rois(100 * 4, 5.0f);
        std::vector<std::int32_t> keep(100, 0), count(1, 0);
        layer.evaluate({MultiArray(probs.data(), {n, 1, 2, 1, 1}), MultiArray(deltas.data(), {n, 1, 4, 1, 1})},
                       {MultiArray(rois.data(), {100, 1, 4, 1, 1})}, keep.data(), count.data());
        save(std::string(tag) + "rois.out", rois);
        save(std::string(tag) + "keep.out", keep);
        save(std::string(tag) + "count.out", count);
      }
      // a layer whose parameters differ from its context must throw at evaluate, not compute something else
      bool threw = false;
      try {
        ProposalLayer other(Parameters{{"maxProposals", std::int64_t(50)}}, ctx);
        std::vector<float> probs(n * 2, 0.5f), rois(50 * 4);
        other.evaluate({MultiArray(probs.data(), {n, 2}), MultiArray(deltas.data(), {n, 4})}, {MultiArray(rois.data(), {50, 4})});
      } catch (const Error& e) {
        threw = e.status() == MRCNN_EINVAL;
      }
      if (!threw) {
        std::fprintf(stderr, "parameter mismatch did not throw\n");
        return 1;
      }
    }
    // ---- default configuration (1024x1024) for the other layers
    MaskRCNNConfig::defaultConfig().maxBatch = 1;
    auto ctx = Context::shared();
    {
      auto rois = load<float>("ra_rois.f32", 64 * 4);
      const std::int64_t side[4] = {32, 16, 8, 4};
      std::vector<std::vector<float>> maps;
      for (int l = 0; l < 4; ++l) maps.push_back(load<float>("maps" + std::to_string(l) + ".f32", static_cast<size_t>(8 * side[l] * side[l])));
      std::vector<MultiArray> in{MultiArray(rois.data(), {64, 1, 4, 1, 1})};
      for (int l = 0; l < 4; ++l) in.emplace_back(maps[static_cast<size_t>(l)].data(), Shape{1, 1, 8, side[l], side[l]});
      std::vector<float> out(64 * 8 * 7 * 7, 3.0f);
      std::vector<std::int32_t> level(64, 0);
      PyramidROIAlignLayer p7(Parameters{{"poolSize", std::int64_t(7)}});
      p7.evaluate(in, {MultiArray(out.data(), {64, 1, 8, 7, 7})}, 1, level.data());
      save("pooled7.out", out);
      save("levels.out", level);
      // detections-shaped rois: rows of 6 (PyramidROIAlignLayer.swift:356 reads the stride from the array)
      std::vector<float> r6(64 * 6, 0.0f);
      for (int i = 0; i < 64; ++i)
        for (int k = 0; k < 4; ++k) r6[static_cast<size_t>(i * 6 + k)] = rois[static_cast<size_t>(i * 4 + k)];
      in[0] = MultiArray(r6.data(), {64, 1, 6, 1, 1});
      std::vector<float> out14(64 * 8 * 14 * 14, 0.0f);
      PyramidROIAlignLayer p14(Parameters{{"poolSize", std::int64_t(14)}});
      p14.evaluate(in, {MultiArray(out14.data(), {64, 1, 8, 14, 14})});
      save("pooled14.out", out14);
    }
    {
      auto probs = load<float>("cls_probs.f32", 200 * 81);
      auto bbox = load<float>("cls_bbox.f32", 200 * 324);
      std::vector<float> sel(200 * 6, 0.0f);
      TimeDistributedClassifierLayer().select(MultiArray(probs.data(), {200, 81}), MultiArray(bbox.data(), {200, 324}),
                                              MultiArray(sel.data(), {200, 6}));
      save("cls_select.out", sel);
      auto rois = load<float>("det_rois.f32", 200 * 4);
      auto cls = load<float>("det_cls.f32", 200 * 6);
      std::vector<float> det(100 * 6, 1.0f);
      std::vector<std::int32_t> keep(100, 0), count(1, 0);
      DetectionLayer layer;
      layer.evaluate({MultiArray(rois.data(), {200, 1, 4, 1, 1}), MultiArray(cls.data(), {200, 1, 1, 1, 6})},
                     {MultiArray(det.data(), {100, 1, 6, 1, 1})}, keep.data(), count.data());
      save("det.out", det);
      save("det_keep.out", keep);
      save("det_count.out", count);
      // Detection.detectionsFromFeatureValue on those detections + the golden masks
      auto masks = load<float>("masks.f32", 100 * 28 * 28);
      MultiArray mv(masks.data(), {100, 28, 28});
      auto dets = Detection::detectionsFromFeatureValue(MultiArray(det.data(), {100, 6}), &mv);
      std::vector<std::int32_t> meta;
      std::vector<double> real;
      std::vector<std::uint8_t> m8;
      for (const Detection& d : dets) {
        meta.push_back(d.index);
        meta.push_back(d.classId);
        real.insert(real.end(), {d.boundingBox.x, d.boundingBox.y, d.boundingBox.width, d.boundingBox.height, d.score});
        if (d.maskSize != 28) return 1;
        m8.insert(m8.end(), d.mask.begin(), d.mask.end());
      }
      save("dec_meta.out", meta);
      save("dec_real.out", real);
      save("dec_mask.out", m8);
      // without a mask feature value the masks stay empty (Detection.swift:47-52)
      auto plain = Detection::detectionsFromFeatureValue(MultiArray(det.data(), {100, 6}), nullptr);
      if (plain.size() != dets.size() || (!plain.empty() && !plain[0].mask.empty())) return 1;
    }
    std::printf("host mirror GPU run ok, kernels launched: %lld\n", static_cast<long long>(ctx->launchCount()));
  } catch (const Error& e) {
    std::fprintf(stderr, "mrcnn::Error %d: %s\n", e.status(), e.what());
    return 1;
  }
  return 0;
}
