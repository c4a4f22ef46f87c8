//  CustomLayers.swift -- the other four MLCustomLayer classes of the reference bound to the C ABI (layer-level drop-in:
//  Core ML still runs the dense graphs of MaskRCNN.mlmodel and calls these classes by their @objc names, as it calls
//  the reference's: Conversion/task.py:27,39,48,53,59).  Same class names, parameter keys and defaults as
//  PyramidROIAlignLayer.swift:40-77, TimeDistributedClassifierLayer.swift:14-32, DetectionLayer.swift:52-105 and
//  TimeDistributedMaskLayer.swift:14-37 of the reference; evaluate() hands the MLMultiArray buffers to the library,
//  which stages host memory through its pinned workspace and always writes every output element.
//  Each layer owns a small context (max_batch = 1; device buffers are allocated on first use).
//  Darwin-only; not compiled in the build container (no Swift toolchain) -- include/maskrcnn.hpp is the same surface
//  in C++ and IS compiled and tested there.
#if canImport(CoreML)
import CoreML
import CMaskRCNNCuda

private func check(_ status: Int32, _ ctx: OpaquePointer?) throws {
    if status != 0 { throw MaskRCNNError(status: status, description: String(cString: mrcnn_last_error(ctx))) }
}

private func floats(_ a: MLMultiArray) -> UnsafeMutablePointer<Float> {
    assert(a.dataType == .float32)
    return a.dataPointer.assumingMemoryBound(to: Float.self)
}

private func readStdDev(_ parameters: [String: Any], into cfg: inout mrcnn_config) {
    // ProposalLayer.swift:70-80 / DetectionLayer.swift:67-77: taken only when all `count` items are Doubles
    guard let n = parameters["bboxStdDev_count"] as? Int, n == 4 else { return }
    var v = [Float]()
    for i in 0..<n { if let d = parameters["bboxStdDev_\(i)"] as? Double { v.append(Float(d)) } }
    if v.count == n { cfg.bbox_std = (v[0], v[1], v[2], v[3]) }
}

private func makeContext(_ cfg: inout mrcnn_config, paths: [String?] = [nil, nil]) throws -> OpaquePointer? {
    var ctx: OpaquePointer?
    cfg.max_batch = 1
    let dup = paths.map { $0.map { strdup($0) } ?? nil }
    defer { dup.forEach { free($0) } }
    cfg.classifier_model_path = UnsafePointer(dup[0])
    cfg.mask_model_path = UnsafePointer(dup[1])
    let status = mrcnn_create(&cfg, &ctx)
    if status != 0 { throw MaskRCNNError(status: status, description: String(cString: mrcnn_last_error(nil))) }
    return ctx
}

@objc(PyramidROIAlignLayer) class PyramidROIAlignLayer: NSObject, MLCustomLayer {
    private var ctx: OpaquePointer?
    private var poolSize = 7                                                                 // PyramidROIAlignLayer.swift:45

    required init(parameters: [String: Any]) throws {
        super.init()
        var cfg = mrcnn_config()
        mrcnn_config_default(&cfg)
        if let v = parameters["poolSize"] as? Int { poolSize = v }                           // :51-53
        // :55-58 reads these `as? CGFloat` although the converter writes Ints (SURVEY.md Q15); both kinds are honoured
        if let w = (parameters["imageWidth"] as? NSNumber)?.intValue,
           let h = (parameters["imageHeight"] as? NSNumber)?.intValue {
            cfg.image_w = Int32(w); cfg.image_h = Int32(h)
        }
        ctx = try makeContext(&cfg)
    }

    deinit { mrcnn_destroy(ctx) }

    func setWeightData(_ weights: [Data]) throws {}

    func outputShapes(forInputShapes inputShapes: [[NSNumber]]) throws -> [[NSNumber]] {     // :65-77
        let rois = inputShapes[0], fmap = inputShapes[1]
        return [[rois[0], rois[1], fmap[2], poolSize as NSNumber, poolSize as NSNumber]]
    }

    func evaluate(inputs: [MLMultiArray], outputs: [MLMultiArray]) throws {                   // :79-181
        let rois = inputs[0]
        let count = Int64(truncating: rois.shape[0])
        let stride = Int32(truncating: rois.strides[0])                                      // 4 or 6 (:356)
        let maps = inputs[1..<5]
        let channels = Int64(truncating: maps.first!.shape[2])
        var pointers = maps.map { UnsafePointer<Float>(floats($0)) as UnsafePointer<Float>? }
        var hw = [Int32]()
        for m in maps { hw.append(Int32(truncating: m.shape[3])); hw.append(Int32(truncating: m.shape[4])) }
        try check(mrcnn_pyramid_roialign_eval(ctx, 1, floats(rois), stride, count, &pointers, &hw, channels,
                                              Int32(poolSize), floats(outputs[0]), nil), ctx)
    }
}

@objc(TimeDistributedClassifierLayer) class TimeDistributedClassifierLayer: NSObject, MLCustomLayer {
    private var ctx: OpaquePointer?

    required init(parameters: [String: Any]) throws {
        super.init()
        var cfg = mrcnn_config()
        mrcnn_config_default(&cfg)
        // the reference loads the Classifier model on every evaluate (TimeDistributedClassifierLayer.swift:41); here once
        ctx = try makeContext(&cfg, paths: [MaskRCNNConfig.defaultConfig.compiledClassifierModelURL!.path, nil])
    }

    deinit { mrcnn_destroy(ctx) }

    func setWeightData(_ weights: [Data]) throws {}

    func outputShapes(forInputShapes inputShapes: [[NSNumber]]) throws -> [[NSNumber]] {     // :26-32
        let s = inputShapes[0]
        return [[s[0], s[1], 1, 1, 6]]
    }

    func evaluate(inputs: [MLMultiArray], outputs: [MLMultiArray]) throws {                   // :34-91
        let count = Int64(truncating: inputs[0].shape[0])
        try check(mrcnn_classifier_eval(ctx, 1, count, floats(inputs[0]), floats(outputs[0])), ctx)
    }
}

@objc(DetectionLayer) class DetectionLayer: NSObject, MLCustomLayer {
    private var ctx: OpaquePointer?
    private var maxDetections = 100                                                          // DetectionLayer.swift:57

    required init(parameters: [String: Any]) throws {
        super.init()
        var cfg = mrcnn_config()
        mrcnn_config_default(&cfg)
        readStdDev(parameters, into: &cfg)                                                   // :67-77
        if let v = parameters["maxDetections"] as? Int { cfg.max_detections = Int32(v); maxDetections = v }      // :79-81
        if let v = parameters["scoreThreshold"] as? Double { cfg.detection_min_score = Float(v) }                 // :82-84
        if let v = parameters["nmsIOUThreshold"] as? Double { cfg.detection_nms_iou = Float(v) }                  // :85-87
        ctx = try makeContext(&cfg)
    }

    deinit { mrcnn_destroy(ctx) }

    func setWeightData(_ weights: [Data]) throws {}

    func outputShapes(forInputShapes inputShapes: [[NSNumber]]) throws -> [[NSNumber]] {     // :94-105
        return [[maxDetections as NSNumber, inputShapes[0][1], 6, 1, 1]]
    }

    func evaluate(inputs: [MLMultiArray], outputs: [MLMultiArray]) throws {                   // :107-234
        let count = Int64(truncating: inputs[0].shape[0])
        try check(mrcnn_detection_eval(ctx, 1, count, floats(inputs[0]), floats(inputs[1]), floats(outputs[0]), nil, nil), ctx)
    }
}

@objc(TimeDistributedMaskLayer) class TimeDistributedMaskLayer: NSObject, MLCustomLayer {
    private var ctx: OpaquePointer?

    required init(parameters: [String: Any]) throws {
        super.init()
        var cfg = mrcnn_config()
        mrcnn_config_default(&cfg)
        ctx = try makeContext(&cfg, paths: [nil, MaskRCNNConfig.defaultConfig.compiledMaskModelURL!.path])      // :49
    }

    deinit { mrcnn_destroy(ctx) }

    func setWeightData(_ weights: [Data]) throws {}

    func outputShapes(forInputShapes inputShapes: [[NSNumber]]) throws -> [[NSNumber]] {     // :26-37
        let s = inputShapes[0]
        return [[1, s[1], s[0], NSNumber(value: s[3].intValue * 2), NSNumber(value: s[4].intValue * 2)]]
    }

    func evaluate(inputs: [MLMultiArray], outputs: [MLMultiArray]) throws {                   // :39-91
        let count = Int64(truncating: inputs[1].shape[0])                                    // detectionCount (:46)
        try check(mrcnn_mask_eval(ctx, 1, count, floats(inputs[0]), floats(inputs[1]), floats(outputs[0])), ctx)
    }
}
#endif
