// swift-tools-version:5.5
// The reference's Package.swift (Package.swift:8-25) declares the library target "Mask-RCNN-CoreML" with no
// dependencies.  This manifest is the same product with the CUDA back end: a system-library target for
// libmaskrcnn_cuda.so (the C ABI of include/maskrcnn_cuda.h) and a Swift target that keeps the reference's public
// types (MaskRCNNConfig, Detection) and adds the MaskRCNN model class.  NOT compiled in the build container (no swift).
import PackageDescription

let package = Package(
    name: "Mask-RCNN-CUDA",
    products: [.library(name: "Mask-RCNN-CoreML", targets: ["MaskRCNNCuda"])],
    targets: [
        .systemLibrary(name: "CMaskRCNNCuda", path: "Sources/CMaskRCNNCuda"),
        .target(name: "MaskRCNNCuda", dependencies: ["CMaskRCNNCuda"], path: "Sources/MaskRCNNCuda"),
    ]
)
