//  ProposalLayer.swift -- how one of the reference's MLCustomLayer classes binds to the C ABI when the rest of the
//  graph still runs in Core ML (layer-level drop-in).  Same @objc name, parameter keys and defaults as
//  ProposalLayer.swift:52-101 of the reference; evaluate() forwards the MLMultiArray buffers to mrcnn_proposal_eval.
//  The other four layers bind the same way (INTEGRATION.md lists the calls).  Darwin-only; not compiled here.
#if canImport(CoreML)
import CoreML
import CMaskRCNNCuda

@objc(ProposalLayer) class ProposalLayer: NSObject, MLCustomLayer {
    private var ctx: OpaquePointer?
    private var maxProposals = 1000

    required init(parameters: [String: Any]) throws {
        super.init()
        var cfg = mrcnn_config()
        mrcnn_config_default(&cfg)
        if let v = parameters["preNMSMaxProposals"] as? Int { cfg.pre_nms_max_proposals = Int32(v) }   // ProposalLayer.swift:82-84
        if let v = parameters["maxProposals"] as? Int { cfg.max_proposals = Int32(v); maxProposals = v } // :85-87
        if let v = parameters["nmsIOUThreshold"] as? Double { cfg.proposal_nms_iou = Float(v) }          // :88-90
        if let n = parameters["bboxStdDev_count"] as? Int, n == 4 {                                       // :70-80
            // all or none, like the reference: the defaults stay unless every one of the `count` items is a Double
            var v = [Float]()
            for i in 0..<n { if let d = parameters["bboxStdDev_\(i)"] as? Double { v.append(Float(d)) } }
            if v.count == n { cfg.bbox_std = (v[0], v[1], v[2], v[3]) }
        }
        // anchors.bin when the configuration names one (:68), else generated for the input size (the reference's own
        // TODO, MaskRCNNConfig.swift:14) -- as MaskRCNN.swift does
        let anchorsPath = MaskRCNNConfig.defaultConfig.anchorsURL?.path
        var status: Int32
        if let path = anchorsPath {
            status = path.withCString { p -> Int32 in cfg.anchors_path = p; return mrcnn_create(&cfg, &ctx) }
        } else {
            cfg.anchors_path = nil
            status = mrcnn_create(&cfg, &ctx)
        }
        if status != 0 { throw MaskRCNNError(status: status, description: String(cString: mrcnn_last_error(nil))) }
        if anchorsPath == nil {
            let n = mrcnn_anchor_count(cfg.image_h, cfg.image_w)
            if n < 0 { throw MaskRCNNError(status: Int32(n), description: "mrcnn_anchor_count: bad image size") }
            var anchors = [Float](repeating: 0, count: Int(n) * 4)
            status = mrcnn_generate_anchors(cfg.image_h, cfg.image_w, &anchors, n)
            if status == 0 { status = mrcnn_set_anchors(ctx, anchors, n) }
            if status != 0 { throw MaskRCNNError(status: status, description: String(cString: mrcnn_last_error(ctx))) }
        }
    }

    deinit { mrcnn_destroy(ctx) }

    func setWeightData(_ weights: [Data]) throws {}                                                       // :93-95

    func outputShapes(forInputShapes inputShapes: [[NSNumber]]) throws -> [[NSNumber]] {                  // :97-101
        var out = inputShapes[1]
        out[0] = NSNumber(value: maxProposals)
        return [out]
    }

    func evaluate(inputs: [MLMultiArray], outputs: [MLMultiArray]) throws {                               // :103-195
        let n = Int64(truncating: inputs[0].shape[0])
        let status = mrcnn_proposal_eval(ctx, 1, n,
                                         inputs[0].dataPointer.assumingMemoryBound(to: Float.self),
                                         inputs[1].dataPointer.assumingMemoryBound(to: Float.self),
                                         outputs[0].dataPointer.assumingMemoryBound(to: Float.self), nil, nil)
        if status != 0 { throw MaskRCNNError(status: status, description: String(cString: mrcnn_last_error(ctx))) }
    }
}
#endif
