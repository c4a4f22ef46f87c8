#!/usr/bin/env python
"""A/B of the per-sub-batch stage launches (MRCNN_STAGE_PARTS, pipeline.cu: all blocks of a ResNet stage for the first
B/p images, then for the next B/p, ..., so that a block's working set fits the L2) against whole-batch launches in ONE
process: one model per variant, the predict calls interleaved, CUDA-event timed; prints median / min ms per batch,
images/s, the launch count per batch and whether the outputs are bit-identical (they must be: tiles never span images).
  VARIANTS="off 1,1,2,1 8,1,1,1 8,4,2,1"   parts per stage res2,res3,res4,res5; "off" = the default path"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m


def main():
    batch = int(os.environ.get("BATCH", "8"))
    rounds = int(os.environ.get("ROUNDS", "12"))
    variants = os.environ.get("VARIANTS", "off 1,1,2,1 8,1,1,1 8,4,2,1").split()
    _, blobs = m.weights.synthetic_blobs(101)
    anchors = m.synth.generate_anchors(1024, 1024)
    img = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (batch, 1024, 1024, 3), dtype=np.uint8)).cuda()
    stream = torch.cuda.Stream()
    models = []
    for v in variants:
        if v != "off":
            os.environ["MRCNN_STAGE_PARTS"] = v
        else:
            os.environ.pop("MRCNN_STAGE_PARTS", None)
        cfg = m.MaskRCNNConfig()
        cfg.maxBatch = batch
        mod = m.MaskRCNN(cfg, blobs=blobs, anchors=anchors)
        mod.ctx.set_stream(stream.cuda_stream)         # one stream for every variant: the events below are recorded on it
        det = torch.zeros((batch, 100, 6), device="cuda"); msk = torch.zeros((batch, 100, 28, 28), device="cuda")
        for _ in range(3):
            mod.prediction_batch(img, det, msk)        # builds the graphs under this variant's environment
        n0 = mod.ctx.launch_count
        mod.prediction_batch(img, det, msk)
        models.append((v, mod, det, msk, mod.ctx.launch_count - n0))
    os.environ.pop("MRCNN_STAGE_PARTS", None)
    torch.cuda.synchronize()
    times = {v: [] for v, *_ in models}
    for _ in range(rounds):
        for v, mod, det, msk, _n in models:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            with torch.cuda.stream(stream):
                e0.record()
                for _ in range(5):
                    mod.prediction_batch(img, det, msk)
                e1.record()
            torch.cuda.synchronize()
            times[v].append(e0.elapsed_time(e1) / 5)
    outs = [(d.cpu().numpy(), k.cpu().numpy()) for _, _, d, k, _ in models]
    same = all(np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1]) for o in outs[1:])
    for v, _, _, _, n in models:
        t = np.array(times[v])
        print(f"parts {v:>8s}: median {np.median(t):.3f} ms  min {t.min():.3f}  -> {batch / np.median(t) * 1e3:.1f} images/s, {n} launches per batch")
    print("outputs identical across variants:", same)
    for _, mod, *_ in models:
        mod.close()
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
