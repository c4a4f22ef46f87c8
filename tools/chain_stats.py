#!/usr/bin/env python
"""Where do the roles of the chained ResNet-stage kernel (conv_chain.cuh) wait?  Runs the batch-8 R101 pipeline, captures
the per-CTA clock64 totals of one predict's chain launches (mrcnn_debug_chain_stats) and prints, per launch, the mean /
max over CTAs of every counter as a fraction of the kernel's duration."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m

NAMES = ["total", "prod:flags", "prod:ring-empty", "mma:operands", "mma:acc-busy", "epi:acc-wait", "epi:residual", "store:bulk-wait",
         "items", "chunks", "flushes", "prod:items-spun", "store:busy", "epi:buffer-wait", "signal:release", "-"]


def main():
    batch = int(os.environ.get("BATCH", "8"))
    cfg = m.MaskRCNNConfig()
    cfg.maxBatch = batch
    _, blobs = m.weights.synthetic_blobs(101)
    model = m.MaskRCNN(cfg, blobs=blobs, anchors=m.synth.generate_anchors(1024, 1024))
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.integers(0, 256, (batch, 1024, 1024, 3), dtype=np.uint8)).cuda()
    det = torch.zeros((batch, 100, 6), device="cuda"); msk = torch.zeros((batch, 100, 28, 28), device="cuda")
    for _ in range(3):
        model.prediction_batch(img, det, msk)
    buf = torch.zeros((8, 148, 16), dtype=torch.int64, device="cuda")
    m._cabi.check(None, m.lib().mrcnn_debug_chain_stats(buf.data_ptr(), 0))
    model.prediction_batch(img, det, msk)
    torch.cuda.synchronize()
    m.lib().mrcnn_debug_chain_stats(None, 0)
    s = buf.cpu().numpy().astype(np.float64)
    for k in range(8):
        a = s[k]
        if a[:, 0].max() == 0:
            continue
        tot = a[:, 0].max()
        print(f"chain launch {k}: {tot:.0f} clk (~{tot / 1.9e3:.0f} us at 1.9 GHz), {int((a[:, 8] > 0).sum())} CTAs, "
              f"{a[:, 8].sum():.0f} CTA items, {a[:, 9].sum():.0f} chunks")
        for i in (1, 2, 3, 4, 5, 6, 7, 12, 13, 14):
            col = a[:, i]
            nz = col[col > 0]
            mean = nz.mean() if len(nz) else 0.0
            print(f"   {NAMES[i]:18s} mean {100 * mean / tot:5.1f}%  max {100 * col.max() / tot:5.1f}%  (threads reporting {len(nz)})")
        print(f"   flushes/CTA {a[:, 10].mean():.1f}  items-spun/CTA {a[:, 11].mean():.1f}")
    model.close()


if __name__ == "__main__":
    main()
