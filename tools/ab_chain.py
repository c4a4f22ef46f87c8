#!/usr/bin/env python
"""A/B of the chained ResNet stages (conv_chain.cuh) against layer-by-layer launches in ONE process: two models, the
predict calls of both interleaved, CUDA-event timed; prints the median / min milliseconds per batch of every variant.
  VARIANTS="0:0 1:4 1:14:0 1:15"  (MRCNN_CHAIN:MRCNN_CHAIN_STAGES[:MRCNN_CHAIN_LAG])"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m


def main():
    batch = int(os.environ.get("BATCH", "8"))
    rounds = int(os.environ.get("ROUNDS", "12"))
    variants = os.environ.get("VARIANTS", "0:0 1:4 1:14").split()
    _, blobs = m.weights.synthetic_blobs(101)
    anchors = m.synth.generate_anchors(1024, 1024)
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.integers(0, 256, (batch, 1024, 1024, 3), dtype=np.uint8)).cuda()
    models = []
    stream = torch.cuda.Stream()
    for v in variants:
        c, mask, *rest = v.split(":")
        os.environ["MRCNN_CHAIN"] = c
        os.environ["MRCNN_CHAIN_STAGES"] = mask
        os.environ["MRCNN_CHAIN_LAG"] = rest[0] if rest else "2"
        for k in ("MRCNN_CHAIN_MAXLEN",):
            os.environ.pop(k, None)
        cfg = m.MaskRCNNConfig()
        cfg.maxBatch = batch
        mod = m.MaskRCNN(cfg, blobs=blobs, anchors=anchors)
        mod.ctx.set_stream(stream.cuda_stream)         # one stream for every variant: the events below are recorded on it
        det = torch.zeros((batch, 100, 6), device="cuda"); msk = torch.zeros((batch, 100, 28, 28), device="cuda")
        for _ in range(3):
            mod.prediction_batch(img, det, msk)        # builds the graphs under this variant's environment
        models.append((v, mod, det, msk))
    torch.cuda.synchronize()
    times = {v: [] for v, *_ in models}
    ref = None
    for r in range(rounds):
        for v, mod, det, msk in models:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            with torch.cuda.stream(stream):
                e0.record()
                for _ in range(5):
                    mod.prediction_batch(img, det, msk)
                e1.record()
            torch.cuda.synchronize()
            times[v].append(e0.elapsed_time(e1) / 5)
    outs = [(d.cpu().numpy(), k.cpu().numpy()) for _, _, d, k in models]
    same = all(np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1]) for o in outs[1:])
    for v in times:
        t = np.array(times[v])
        print(f"variant {v:>6s}: median {np.median(t):.3f} ms  min {t.min():.3f}  -> {batch / np.median(t) * 1e3:.1f} images/s")
    print("outputs identical across variants:", same)
    for _, mod, _, _ in models:
        mod.close()


if __name__ == "__main__":
    main()
