#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ from the CPU oracle (oracle/oracle.c).

The reference (Swift + Core ML + Accelerate + MPS) ships no golden vectors and cannot run on Linux, so these fixtures
pin the ORACLE's behaviour (and through it the CUDA path) at the time they were generated; they are small seeded
cases plus the adversarial set of SURVEY.md section 8(d)(vi).  Re-run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import maskrcnn_b200 as m            # noqa: E402  (synthetic input generators only; no GPU needed)
from oracle import oracle as orc     # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SIZE = 128


def main():
    anchors = m.synth.generate_anchors(SIZE, SIZE)                       # 4092 anchors
    # ---- ProposalLayer: clustered RPN outputs, heavy score ties, and < max survivors
    probs, deltas = m.synth.rpn_outputs(anchors, 3, image_size=SIZE)
    rois, keep, cnt = orc.proposal(probs, deltas, anchors, pre_nms=600, max_proposals=100)
    tp = probs.copy()
    tp[:, 1] = np.round(tp[:, 1] * 8) / 8
    tp[:, 0] = 1 - tp[:, 1]
    trois, tkeep, tcnt = orc.proposal(tp, deltas, anchors, pre_nms=600, max_proposals=100)
    np.savez_compressed(os.path.join(OUT, "proposal.npz"), anchors=anchors, probs=probs, deltas=deltas, rois=rois, keep=keep, count=cnt,
                        tie_probs=tp, tie_rois=trois, tie_keep=tkeep, tie_count=tcnt)
    # ---- PyramidROIAlign: all four levels, padding rois, x.5 level boundaries, pool 7 and 14
    maps = m.synth.feature_maps(5, SIZE, SIZE, channels=8)
    r = m.synth.random_rois(64, 5, min_px=4, max_px=120, image_size=SIZE, n_pad=4)
    r[10] = [0.25, 0.25, 0.25, 0.75]                                       # zero height -> padding block
    ratio = 224.0 / 1024.0      # the layer is configured for a 1024x1024 image here so that all four levels occur
    for i, lv in enumerate((2.5, 3.5, 4.5)):                               # exact half levels (round half away from zero, Q13)
        side = ratio * 2.0 ** (lv - 4.0)
        r[20 + i] = [0.0, 0.0, min(side, 1.0), min(side, 1.0)]
    p7, l7 = orc.pyramid_roialign(r, maps, 7, 1024, 1024)
    r6 = np.concatenate([r, np.zeros((64, 2), np.float32)], axis=1)
    p14, l14 = orc.pyramid_roialign(r6, maps, 14, 1024, 1024)
    np.savez_compressed(os.path.join(OUT, "roialign.npz"), rois=r, maps0=maps[0], maps1=maps[1], maps2=maps[2], maps3=maps[3],
                        pooled7=p7, levels=l7, pooled14=p14)
    # ---- classifier select + DetectionLayer (clustered same-class boxes, threshold edge, ties)
    pr, bb = m.synth.classifier_outputs(200, 9)
    pr[3, 5] = pr[3, 9] = 0.99
    cls = orc.classifier_select(pr, bb)
    cls[:, :4] *= 0.5
    dr = m.synth.random_rois(200, 9, min_px=10, max_px=100, image_size=SIZE)
    dr[50:120] = np.clip(dr[50] + np.random.default_rng(1).uniform(-0.02, 0.02, (70, 4)).astype(np.float32), 0, 1)
    cls[50:120, 4] = 7.0
    cls[50:120, 5] = np.float32(0.9)
    cls[cls[:, 4] >= 78, 4] = 1.0
    cls[130, 4], cls[131, 4] = 78.0, 79.0                                 # classes of their own: never suppressed
    cls[130, 5] = np.float32(0.7)
    cls[131, 5] = np.nextafter(np.float32(0.7), np.float32(0))
    det, dkeep, dcnt = orc.detection(dr, cls)
    np.savez_compressed(os.path.join(OUT, "detection.npz"), probs=pr, bbox=bb, cls=cls, rois=dr, det=det, keep=dkeep, count=dcnt)
    # ---- mask class-plane selection + public Detection decoding
    rng = np.random.default_rng(4)
    masks_all = rng.uniform(size=(100, 81, 4, 4)).astype(np.float32)
    valid = (np.arange(100) < dcnt).astype(np.int32)
    sel = orc.mask_select(masks_all, valid, det)
    big = rng.uniform(size=(100, 28, 28)).astype(np.float32)
    n, idx, bbox, dc, score, mu8 = orc.detections_decode(det, big)
    np.savez_compressed(os.path.join(OUT, "mask_decode.npz"), masks_all=masks_all, valid=valid, selected=sel, masks=big, n=n,
                        index=idx, bbox=bbox, classes=dc, score=score, mask_u8=mu8)
    print("golden fixtures written:", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
