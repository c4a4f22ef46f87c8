#!/usr/bin/env python
"""ROIAlign-only microbench (BASELINE.json configs[4]): R rois constrained to pyramid level 2 (P2 = 256x256x256) per
image; sweep R x batch x pool x layout.  HBM GB/s uses the compulsory-traffic model of SURVEY.md 8(d)
(every touched map read once + output written once + rois); the tap model is printed as "effective".
L2 is flushed (256 MB memset) before every timed launch."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def level2_rois(n, seed):
    rng = np.random.default_rng(seed)
    side = np.exp(rng.uniform(np.log(12.0), np.log(75.0), n))          # sqrt(w*h) < 79.2 px -> level 2
    aspect = np.exp(rng.uniform(np.log(0.6), np.log(1.6), n))
    h, w = side / np.sqrt(aspect) / 1024.0, side * np.sqrt(aspect) / 1024.0
    y1, x1 = rng.uniform(0, 1 - h), rng.uniform(0, 1 - w)
    return np.stack([y1, x1, y1 + h, x1 + w], 1).astype(np.float32)


def main():
    ctx = m.Context()
    st = torch.cuda.Stream()
    ctx.set_stream(st.cuda_stream)
    lib = m.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    hw = (C.c_int32 * 8)(256, 256, 128, 128, 64, 64, 32, 32)
    print(f"{'layout':10s} {'batch':>5s} {'R':>5s} {'pool':>4s} {'us':>9s} {'GB/s':>8s} {'frac':>6s} {'eff GB/s (taps)':>16s}")
    rows = []
    for layout in ("nhwc_f16", "chw_f32"):
        for batch in (1, 8, 64):
            es = 2 if layout == "nhwc_f16" else 4
            dt = torch.float16 if layout == "nhwc_f16" else torch.float32
            shape = lambda s: (batch, s, s, 256) if layout == "nhwc_f16" else (batch, 256, s, s)
            if batch == 64 and layout == "chw_f32":
                continue                                                  # 64 x 89 MB fp32 pyramids + outputs: skipped (memory / time)
            maps = [torch.randn(shape(s), device="cuda", dtype=dt) for s in (256, 128, 64, 32)]
            fp = (C.c_void_p * 4)(*[t.data_ptr() for t in maps])
            for r in (100, 300, 1000):
                rois = torch.from_numpy(np.stack([level2_rois(r, 100 * b + r) for b in range(batch)])).cuda()
                for pool in (7, 14):
                    oshape = (batch, r, pool, pool, 256) if layout == "nhwc_f16" else (batch, r, 256, pool, pool)
                    out = torch.empty(oshape, device="cuda", dtype=dt)
                    ts = []
                    with torch.cuda.stream(st):
                        for it in range(6):
                            flush.zero_()
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record()
                            if layout == "nhwc_f16":
                                rc = lib.mrcnn_roialign_nhwc_f16(ctx.handle, batch, rois.data_ptr(), 4, r, fp, hw, 256, pool, out.data_ptr(), None)
                            else:
                                rc = lib.mrcnn_pyramid_roialign_eval(ctx.handle, batch, rois.data_ptr(), 4, r, fp, hw, 256, pool, out.data_ptr(), None)
                            e1.record()
                            m._cabi.check(ctx.handle, rc)
                            st.synchronize()
                            if it >= 2:
                                ts.append(e0.elapsed_time(e1))
                    ms = float(np.median(ts))
                    byts = batch * (256 * 256 * 256 * es + r * 256 * pool * pool * es + r * 16)   # P2 once + out once + rois
                    taps = batch * r * 256 * pool * pool * (4 * es + es)
                    gbs = byts / ms / 1e6
                    print(f"{layout:10s} {batch:5d} {r:5d} {pool:4d} {ms*1e3:9.1f} {gbs:8.0f} {gbs/PEAK:6.2f} {taps/ms/1e6:16.0f}")
                    rows.append({"layout": layout, "batch": batch, "rois": r, "pool": pool, "us": ms * 1e3, "GBps": gbs, "frac_of_measured_hbm": gbs / PEAK})
                    del out
            del maps
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"peak_hbm_GBps": PEAK, "includes_level_kernel": True, "rows": rows}, open("gpurun_out/roialign_sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()
