#!/usr/bin/env python
"""Converts the reference's model split (MaskRCNN.mlmodel, Classifier.mlmodel, Mask.mlmodel: README.md:107-116) into the
MRCNNW1 blobs libmaskrcnn_cuda.so loads (mrcnn_config.{main,classifier,mask}_model_path) and, on request, writes
anchors.bin next to them (the reference's own TODO: "generate the anchors on demand", MaskRCNNConfig.swift:14).
  python tools/import_mlmodel.py --main MaskRCNN.mlmodel --classifier Classifier.mlmodel --mask Mask.mlmodel --out products/"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--main", required=True)
    ap.add_argument("--classifier", required=True)
    ap.add_argument("--mask", required=True)
    ap.add_argument("--architecture", type=int, default=101, choices=[50, 101])
    ap.add_argument("--num-classes", type=int, default=81)
    ap.add_argument("--out", required=True)
    ap.add_argument("--anchors", action="store_true", help="also write anchors.bin for the model's input size")
    a = ap.parse_args()
    _, blobs, extra = m.mlmodel.import_products(a.main, a.classifier, a.mask, a.architecture, a.num_classes)
    os.makedirs(a.out, exist_ok=True)
    for name, blob in zip(("MaskRCNN", "Classifier", "Mask"), blobs):
        with open(os.path.join(a.out, name + ".mrcnnw"), "wb") as f:
            f.write(blob)
        print(f"{name}.mrcnnw: {len(blob)} bytes")
    cfg = m.mlmodel.config_from_custom_layers(extra["custom_layers"], m.MaskRCNNConfig())
    print("custom-layer parameters:", extra["custom_layers"])
    print("mean_rgb:", extra.get("mean_rgb"))
    if a.anchors:
        h, w = cfg.imageShape[:2]
        m.synth.generate_anchors(h, w).tofile(os.path.join(a.out, "anchors.bin"))
        print(f"anchors.bin for {h}x{w}")


if __name__ == "__main__":
    main()
