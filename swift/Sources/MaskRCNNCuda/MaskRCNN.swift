//  MaskRCNN.swift -- Swift host shim over the C ABI of libmaskrcnn_cuda.so.
//
//  Keeps the public surface of Sources/Mask-RCNN-CoreML in the reference:
//    MaskRCNNConfig.defaultConfig + anchorsURL / compiledClassifierModelURL / compiledMaskModelURL
//                                                        (MaskRCNNConfig.swift:10-18)
//    Detection { index, boundingBox, classId, score, mask }   (Detection.swift:15-21)
//  and adds the model class that Xcode generates from MaskRCNN.mlmodel in the reference
//  (Example/Source/ViewController.swift:37): MaskRCNN().predict(image:).
//  This file is not compiled in the build container (no Swift toolchain); tests drive the same C calls via ctypes.
import Foundation
import CMaskRCNNCuda

public struct MaskRCNNError: Error, CustomStringConvertible {
    public let status: Int32
    public let description: String
}

public class MaskRCNNConfig {
    public static let defaultConfig = MaskRCNNConfig()
    public var anchorsURL: URL?
    public var compiledClassifierModelURL: URL?
    public var compiledMaskModelURL: URL?
    public var modelURL: URL?                       // the MaskRCNN model itself (a bundle resource in the reference)
    // layer parameters the reference bakes into the .mlmodel (ProposalLayer.swift:57-63, DetectionLayer.swift:55-61)
    public var architecture: Int32 = 101
    public var imageSize: (height: Int32, width: Int32) = (1024, 1024)
    public var preNMSMaxProposals: Int32 = 6000
    public var maxProposals: Int32 = 1000
    public var maxDetections: Int32 = 100
    public var maxBatch: Int32 = 8
    public var preciseMasks = true                  // mrcnn_config.precise_masks: 2-term fp16 activations in the mask head (masks within 1e-4 of fp32)
}

public struct Detection {
    public let index: Int
    public let boundingBox: CGRect      // normalised (x, y, w, h), Detection.swift:41-55
    public let classId: Int
    public let score: Double
    public let mask: [UInt8]?           // 28x28, 255 - p/2*255 (Detection.swift:83-85)
}

public final class MaskRCNN {
    private var ctx: OpaquePointer?
    private let maxDetections: Int
    private let maskSize: Int
    private let imageBytes: Int

    public init(configuration: MaskRCNNConfig = .defaultConfig) throws {
        var cfg = mrcnn_config()
        mrcnn_config_default(&cfg)
        cfg.architecture = configuration.architecture
        cfg.image_h = configuration.imageSize.height
        cfg.image_w = configuration.imageSize.width
        cfg.pre_nms_max_proposals = configuration.preNMSMaxProposals
        cfg.max_proposals = configuration.maxProposals
        cfg.max_detections = configuration.maxDetections
        cfg.max_batch = configuration.maxBatch
        cfg.precise_masks = configuration.preciseMasks ? 1 : 0
        maxDetections = Int(cfg.max_detections)
        maskSize = 2 * Int(cfg.pool_size_mask)
        imageBytes = Int(cfg.image_h) * Int(cfg.image_w) * 3
        // the C strings must outlive mrcnn_create only
        let paths = [configuration.anchorsURL, configuration.modelURL,
                     configuration.compiledClassifierModelURL, configuration.compiledMaskModelURL].map { $0?.path }
        let status: Int32 = withCStrings(paths) { p in
            cfg.anchors_path = p[0]; cfg.main_model_path = p[1]
            cfg.classifier_model_path = p[2]; cfg.mask_model_path = p[3]
            return mrcnn_create(&cfg, &ctx)
        }
        if status != 0 { throw MaskRCNNError(status: status, description: String(cString: mrcnn_last_error(nil))) }
        if configuration.anchorsURL == nil {
            // the reference's own TODO (MaskRCNNConfig.swift:14): anchors generated for the input size instead of anchors.bin
            let n = mrcnn_anchor_count(cfg.image_h, cfg.image_w)
            if n < 0 { throw MaskRCNNError(status: Int32(n), description: "mrcnn_anchor_count: bad image size") }
            var anchors = [Float](repeating: 0, count: Int(n) * 4)
            var st = mrcnn_generate_anchors(cfg.image_h, cfg.image_w, &anchors, n)
            if st == 0 { st = mrcnn_set_anchors(ctx, anchors, n) }
            if st != 0 { throw MaskRCNNError(status: st, description: String(cString: mrcnn_last_error(ctx))) }
        }
    }

    deinit { mrcnn_destroy(ctx) }

    /// images: B tightly packed RGB8 images already letter-boxed to the model size (Vision's .scaleFit,
    /// EvaluateCommand.swift:157).  Returns the raw model outputs "detections" and "mask" (Conversion/task.py:70-72).
    public func prediction(images: UnsafePointer<UInt8>, batch: Int) throws -> (detections: [Float], masks: [Float]) {
        var det = [Float](repeating: 0, count: batch * maxDetections * 6)
        var msk = [Float](repeating: 0, count: batch * maxDetections * maskSize * maskSize)
        let status = mrcnn_predict(ctx, Int32(batch), images, &det, &msk)
        if status != 0 { throw MaskRCNNError(status: status, description: String(cString: mrcnn_last_error(ctx))) }
        return (det, msk)
    }

    /// Streaming: the loop of EvaluateCommand.swift:166-194 with two batches in flight.  `submit` enqueues a batch and
    /// returns at once (the host->device copy of this batch overlaps the compute of the previous one); `wait` blocks
    /// until the OLDEST submitted batch is complete.  The buffers (ideally page-locked) must stay valid and untouched
    /// until the matching `wait` returns.  At most two batches may be in flight.
    public func submit(images: UnsafePointer<UInt8>, batch: Int, detections: UnsafeMutablePointer<Float>,
                       masks: UnsafeMutablePointer<Float>) throws {
        let status = mrcnn_predict_submit(ctx, Int32(batch), images, detections, masks, 0)
        if status != 0 { throw MaskRCNNError(status: status, description: String(cString: mrcnn_last_error(ctx))) }
    }

    public func wait() throws {
        let status = mrcnn_predict_wait(ctx)
        if status != 0 { throw MaskRCNNError(status: status, description: String(cString: mrcnn_last_error(ctx))) }
    }

    public var batchesInFlight: Int { return Int(mrcnn_predict_in_flight(ctx)) }

    /// One image -> [Detection] with score > 0.7, as ViewController.swift:163-187 does with the Core ML outputs.
    public func predict(image: UnsafePointer<UInt8>) throws -> [Detection] {
        let out = try prediction(images: image, batch: 1)
        let d = maxDetections, s = maskSize
        var count: Int32 = 0
        var index = [Int32](repeating: 0, count: d), cls = [Int32](repeating: 0, count: d)
        var bbox = [Double](repeating: 0, count: 4 * d), score = [Double](repeating: 0, count: d)
        var mask = [UInt8](repeating: 0, count: d * s * s)
        let status = mrcnn_detections_decode(ctx, 1, out.detections, out.masks, &count, &index, &bbox, &cls, &score, &mask)
        if status != 0 { throw MaskRCNNError(status: status, description: String(cString: mrcnn_last_error(ctx))) }
        return (0..<Int(count)).map { i in
            Detection(index: Int(index[i]),
                      boundingBox: CGRect(x: bbox[4 * i], y: bbox[4 * i + 1], width: bbox[4 * i + 2], height: bbox[4 * i + 3]),
                      classId: Int(cls[i]), score: score[i], mask: Array(mask[(i * s * s)..<((i + 1) * s * s)]))
        }
    }
}

private func withCStrings<R>(_ strings: [String?], _ body: ([UnsafePointer<CChar>?]) -> R) -> R {
    let dup = strings.map { $0.map { strdup($0) } ?? nil }
    defer { dup.forEach { free($0) } }
    return body(dup.map { UnsafePointer($0) })
}
