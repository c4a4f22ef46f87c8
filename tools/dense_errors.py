#!/usr/bin/env python
"""Measured numerical distance between the tensor-core dense stages and the PyTorch fp32 restatement (oracle/dense_ref.py),
ResNet50 / 256x256 (the test configuration) and ResNet101 / 1024x1024 batch 1: with fp16 rounding at the same points
(accumulation-order differences only) and against a pure fp32 reference."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m
from oracle.dense_ref import Ref


def run(arch, size, props):
    folded, blobs = m.weights.synthetic_blobs(arch)
    cfg = m.MaskRCNNConfig()
    cfg.architecture, cfg.imageShape, cfg.maxProposals, cfg.maxBatch = f"resnet{arch}", (size, size, 3), props, 1
    model = m.MaskRCNN(cfg, blobs=blobs, anchors=m.synth.generate_anchors(size, size))
    rng = np.random.default_rng(20260)
    img = rng.integers(0, 256, (1, size // 8, size // 8, 3)).astype(np.uint8).repeat(8, 1).repeat(8, 2)
    sizes = [(size // s, size // s) for s in (4, 8, 16, 32)]
    fm = [torch.zeros((1, h, w, 256), dtype=torch.float16, device="cuda") for h, w in sizes]
    n = sum(3 * (size // s) ** 2 for s in (4, 8, 16, 32, 64))
    probs = torch.zeros((1, n, 2), device="cuda"); deltas = torch.zeros((1, n, 4), device="cuda")
    fp = (C.c_void_p * 4)(*[t.data_ptr() for t in fm])
    m._cabi.check(model.ctx.handle, m.lib().mrcnn_backbone_eval(model.ctx.handle, 1, m._cabi.ptr(img), fp, probs.data_ptr(), deltas.data_ptr()))
    for half in (True, False):
        ref = Ref(folded, arch, act_half=half)
        rfm, rp, rd = ref.backbone(img)
        tag = "fp16-rounded reference" if half else "pure fp32 reference"
        for l in range(4):
            d = (fm[l].float() - rfm[l]).abs()
            print(f"  resnet{arch}@{size} P{l+2} vs {tag}: max |d| / max|ref| = {float(d.max() / rfm[l].abs().max()):.2e}, mean |d| / mean|ref| = {float(d.mean() / rfm[l].abs().mean()):.2e}")
        print(f"  resnet{arch}@{size} rpn probs vs {tag}: max |d| = {float((probs - rp).abs().max()):.2e}; deltas max |d| / max|ref| = {float((deltas - rd).abs().max() / rd.abs().max()):.2e}")
        pooled = rng.standard_normal((64, 14, 14, 256)).astype(np.float16)
        det = np.zeros((1, 100, 6), np.float32); det[0, :64, 4] = rng.integers(1, 81, 64); det[0, :64, 5] = 0.9
        chw = np.zeros((1, 100, 256, 14, 14), np.float32); chw[0, :64] = pooled.astype(np.float32).transpose(0, 3, 1, 2)
        out = np.zeros((1, 100, 28, 28), np.float32)
        m.TimeDistributedMaskLayer(context=model.ctx).evaluate([chw, det], [out])
        want = ref.mask(pooled).cpu().numpy()[np.arange(64), det[0, :64, 4].astype(int)]
        print(f"  mask head vs {tag}: max |d| = {np.abs(out[0, :64] - want).max():.2e}")
        p7 = rng.standard_normal((props, 7, 7, 256)).astype(np.float16)
        chw7 = np.ascontiguousarray(p7.astype(np.float32).transpose(0, 3, 1, 2))[None]
        o6 = np.zeros((1, props, 6), np.float32)
        m.TimeDistributedClassifierLayer(context=model.ctx).evaluate([chw7], [o6])
        pr, bb, _ = ref.classifier(p7)
        pr = pr.cpu().numpy(); cls = o6[0, :, 4].astype(int)
        agree = (cls == pr.argmax(1)).mean()
        print(f"  classifier head vs {tag}: argmax agreement {agree:.4f}, max |score d| = {np.abs(o6[0, :, 5] - pr[np.arange(props), cls]).max():.2e}")
    model.close()


if __name__ == "__main__":
    run(50, 256, 200)
    run(101, 1024, 1000)
