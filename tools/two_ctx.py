#!/usr/bin/env python
"""Throughput of TWO contexts (own streams, own activation buffers) predicting alternate batches concurrently, against
one context predicting the same number of batches: do the kernels of one batch fill the SMs the other leaves idle
(partial last rounds, launch heads / tails)?  Device-resident inputs, wall clock around a synchronised region."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m


def main():
    batch = int(os.environ.get("BATCH", "8"))
    n = int(os.environ.get("N", "40"))
    nctx = int(os.environ.get("NCTX", "2"))
    _, blobs = m.weights.synthetic_blobs(101)
    anchors = m.synth.generate_anchors(1024, 1024)
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.integers(0, 256, (batch, 1024, 1024, 3), dtype=np.uint8)).cuda()
    mods = []
    for _ in range(nctx):
        cfg = m.MaskRCNNConfig(); cfg.maxBatch = batch
        mod = m.MaskRCNN(cfg, blobs=blobs, anchors=anchors)
        det = torch.zeros((batch, 100, 6), device="cuda"); msk = torch.zeros((batch, 100, 28, 28), device="cuda")
        for _ in range(3):
            mod.prediction_batch(img, det, msk)
        mods.append((mod, det, msk))
    torch.cuda.synchronize()

    def run(k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            mod, det, msk = mods[i % k]
            mod.prediction_batch(img, det, msk)          # device pointers: returns after enqueue
        for mod, _, _ in mods[:k]:
            mod.ctx.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    for rep in range(3):
        for k in range(1, nctx + 1):
            ms = run(k)
            print(f"rep {rep}: {k} context(s): {ms:.3f} ms / batch -> {batch / ms * 1e3:.1f} images/s")
    ref = mods[0][1].cpu().numpy()
    print("same detections from every context:", all(np.array_equal(ref, d.cpu().numpy()) for _, d, _ in mods[1:]))
    for mod, _, _ in mods:
        mod.close()


if __name__ == "__main__":
    main()
