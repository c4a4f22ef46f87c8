#!/usr/bin/env python
"""Converts Matterport's Keras weight file (mask_rcnn_coco.h5, the input of the reference's conversion task,
Conversion/task.py:163) into the MRCNNW1 blobs libmaskrcnn_cuda.so loads, without Keras / h5py (h5lite reads the HDF5
structures directly), and on request writes anchors.bin for the model's input size.
  python tools/import_keras_h5.py --weights mask_rcnn_coco.h5 --out products/ --anchors"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--weights", required=True)
    ap.add_argument("--architecture", type=int, default=101, choices=[50, 101])
    ap.add_argument("--num-classes", type=int, default=81)
    ap.add_argument("--image-size", type=int, default=1024)
    ap.add_argument("--out", required=True)
    ap.add_argument("--anchors", action="store_true", help="also write anchors.bin for --image-size")
    a = ap.parse_args()
    blobs = m.keras_h5.import_products(a.weights, a.architecture, a.num_classes)
    os.makedirs(a.out, exist_ok=True)
    for name, blob in zip(("MaskRCNN", "Classifier", "Mask"), blobs):
        with open(os.path.join(a.out, name + ".mrcnnw"), "wb") as f:
            f.write(blob)
        print(f"{name}.mrcnnw: {len(blob)} bytes")
    if a.anchors:
        m.synth.generate_anchors(a.image_size, a.image_size).tofile(os.path.join(a.out, "anchors.bin"))
        print(f"anchors.bin for {a.image_size}x{a.image_size}")


if __name__ == "__main__":
    main()
