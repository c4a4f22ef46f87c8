#!/usr/bin/env python
"""ROIAlign-only microbench (BASELINE.json configs[4]; `bench.py --workload roialign` runs `sweep()` below):
R rois constrained to pyramid level 2 (P2 = 256x256x256) per image; sweep R x batch x pool x layout, L2 flushed
(256 MB memset) before every timed launch.

Three byte counts per case, all per launch:
  touched  : what the launch MUST move -- the map bytes its rois really touch (32-byte sectors of the distinct tap
             pixels, computed here from the roi footprints) + the output once + the rois.  `frac` uses this one.
  survey   : SURVEY.md 8(d)'s compulsory model -- the whole P2 map once + output once + rois (charges pixels no roi
             reads; an upper bound of `touched`, reported for continuity with round 1).
  dram     : dram__bytes_read.sum + dram__bytes_write.sum of the same launch from a committed ncu capture
             (profiles/r2_roialign_dram.json), when one exists for the case.
Times: `us_kernel` = the ROIAlign kernel alone (the library's event pair around that launch), `us_call` = the
C-ABI call (level kernel + ROIAlign kernel), CUDA events on the library's stream.

    python tools/bench_roialign.py                      # full sweep
    python tools/bench_roialign.py --case nhwc_f16,8,1000,7 --iters 3     # one case (for ncu)
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

P2 = 256          # P2 side (pixels) and channels of configs[4]
CH = 256


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return (json.load(open(p))["hbm_gbs"], "measured") if os.path.exists(p) else (6650.0, "fallback")


def ncu_dram_bytes(case_key):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_roialign_dram.json")))[case_key]["dram_bytes"]
    except (OSError, KeyError, ValueError):
        return None


def level2_rois(n, seed):
    rng = np.random.default_rng(seed)
    side = np.exp(rng.uniform(np.log(12.0), np.log(75.0), n))          # sqrt(w*h) < 79.2 px -> level 2
    aspect = np.exp(rng.uniform(np.log(0.6), np.log(1.6), n))
    h, w = side / np.sqrt(aspect) / 1024.0, side * np.sqrt(aspect) / 1024.0
    y1, x1 = rng.uniform(0, 1 - h), rng.uniform(0, 1 - w)
    return np.stack([y1, x1, y1 + h, x1 + w], 1).astype(np.float32)


def axis_taps(a1, a2, D, P):
    """floor / ceil taps of the P samples along one axis (fp32, operation by operation, as the kernels and the oracle)."""
    dm1 = np.float32(D - 1)
    scale = ((a2 - a1) * dm1) / np.float32(P - 1)
    inn = (a1 * dm1)[:, None] + np.arange(P, dtype=np.float32)[None, :] * scale[:, None]
    ok = ~((inn < 0) | (inn > dm1))
    return np.floor(inn).astype(np.int64), np.ceil(inn).astype(np.int64), ok


def touched_map_bytes(rois, H, W, C, P, layout):
    """Bytes of ONE image's map that its rois read at least once, at 32-byte sector granularity."""
    ylo, yhi, yok = axis_taps(rois[:, 0], rois[:, 2], H, P)
    xlo, xhi, xok = axis_taps(rois[:, 1], rois[:, 3], W, P)
    mask = np.zeros((H, W), bool)
    for i in range(len(rois)):
        ys = np.unique(np.concatenate([ylo[i][yok[i]], yhi[i][yok[i]]]))
        xs = np.unique(np.concatenate([xlo[i][xok[i]], xhi[i][xok[i]]]))
        if len(ys) and len(xs):
            mask[np.ix_(ys, xs)] = True
    if layout == "nhwc_f16":
        return int(mask.sum()) * C * 2                     # a pixel = C * 2 contiguous bytes (16 sectors at C = 256)
    sect = mask.reshape(H, W // 8, 8).any(axis=2)          # CHW fp32: 8 pixels of one row per sector
    return int(sect.sum()) * 32 * C


def run_case(m, torch, ctx, st, flush, layout, batch, r, pool, iters=6, maps=None):
    lib = m.lib()
    es = 2 if layout == "nhwc_f16" else 4
    dt = torch.float16 if layout == "nhwc_f16" else torch.float32
    hw = (C.c_int32 * 8)(256, 256, 128, 128, 64, 64, 32, 32)
    if maps is None:
        maps = make_maps(torch, layout, batch)
    fp = (C.c_void_p * 4)(*[t.data_ptr() for t in maps])
    rois_np = np.stack([level2_rois(r, 100 * b + r) for b in range(batch)])
    rois = torch.from_numpy(rois_np).cuda()
    oshape = (batch, r, pool, pool, CH) if layout == "nhwc_f16" else (batch, r, CH, pool, pool)
    out = torch.empty(oshape, device="cuda", dtype=dt)
    call, kern = [], []
    with torch.cuda.stream(st):
        for it in range(iters + 2):
            flush.zero_()
            ctx.profile_enable(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if layout == "nhwc_f16":
                rc = lib.mrcnn_roialign_nhwc_f16(ctx.handle, batch, rois.data_ptr(), 4, r, fp, hw, CH, pool, out.data_ptr(), None)
            else:
                rc = lib.mrcnn_pyramid_roialign_eval(ctx.handle, batch, rois.data_ptr(), 4, r, fp, hw, CH, pool, out.data_ptr(), None)
            e1.record()
            m._cabi.check(ctx.handle, rc)
            st.synchronize()
            prof = ctx.profile_read()
            ctx.profile_enable(False)
            if it >= 2:
                call.append(e0.elapsed_time(e1) * 1e3)
                kern.append(prof["roialign"][0] * 1e3)
    us_call, us_kernel = float(np.median(call)), float(np.median(kern))
    out_bytes = batch * r * CH * pool * pool * es + batch * r * 16
    touched = sum(touched_map_bytes(rois_np[b], P2, P2, CH, pool, layout) for b in range(batch)) + out_bytes
    survey = batch * P2 * P2 * CH * es + out_bytes
    taps = batch * r * CH * pool * pool * (4 * es + es)
    peak, _ = hbm_peak()
    key = f"{layout},{batch},{r},{pool}"
    dram = ncu_dram_bytes(key)
    gbs = lambda b: b / us_kernel / 1e3
    row = {"case": key, "layout": layout, "batch": batch, "rois": r, "pool": pool, "us_kernel": us_kernel, "us_call": us_call,
           "bytes_touched": touched, "bytes_survey": survey, "bytes_dram_ncu": dram,
           "GBps_touched": gbs(touched), "frac": gbs(touched) / peak,
           "GBps_survey": gbs(survey), "frac_survey": gbs(survey) / peak,
           "GBps_dram_ncu": gbs(dram) if dram else None, "frac_dram_ncu": gbs(dram) / peak if dram else None,
           "GBps_taps_effective": gbs(taps)}
    del out
    return row


def make_maps(torch, layout, batch):
    dt = torch.float16 if layout == "nhwc_f16" else torch.float32
    shape = (lambda s: (batch, s, s, CH)) if layout == "nhwc_f16" else (lambda s: (batch, CH, s, s))
    # only P2 is read (all rois are level 2); the other levels are small placeholders of the right shape
    return [torch.randn(shape(s), device="cuda", dtype=dt) for s in (256, 128, 64, 32)]


def sweep(m, torch, cases=None, iters=6, log=None):
    ctx = m.Context()
    st = torch.cuda.Stream()
    ctx.set_stream(st.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    if cases is None:
        cases = [(layout, batch, r, pool) for layout in ("nhwc_f16", "chw_f32") for batch in (1, 8, 64)
                 for r in (100, 300, 1000) for pool in (7, 14)
                 if not (layout == "chw_f32" and batch == 64)]      # 64 x 67 MB fp32 maps + 3.2 GB outputs: skipped
    cur, maps = None, None
    for layout, batch, r, pool in cases:
        if (layout, batch) != cur:
            del maps
            maps = make_maps(torch, layout, batch)
            cur = (layout, batch)
        row = run_case(m, torch, ctx, st, flush, layout, batch, r, pool, iters, maps)
        rows.append(row)
        if log:
            log(f"{layout:9s} b{batch:<3d} R{r:<5d} P{pool:<3d} kernel {row['us_kernel']:8.1f} us  call {row['us_call']:8.1f} us  "
                f"touched {row['GBps_touched']:6.0f} GB/s ({row['frac']:.2f})  survey {row['GBps_survey']:6.0f} ({row['frac_survey']:.2f})"
                + (f"  dram(ncu) {row['GBps_dram_ncu']:6.0f} ({row['frac_dram_ncu']:.2f})" if row["bytes_dram_ncu"] else "")
                + f"  taps {row['GBps_taps_effective']:6.0f}")
    ctx.close()
    return rows


def main():
    import torch
    import maskrcnn_b200 as m
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None, help="layout,batch,rois,pool (one case; for ncu)")
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--out", default="gpurun_out/roialign_sweep.json")
    args = ap.parse_args()
    cases = None
    if args.case:
        cases = []
        for c in args.case.split(";"):
            l, b, r, p = c.split(",")
            cases.append((l, int(b), int(r), int(p)))
    rows = sweep(m, torch, cases, args.iters, log=lambda s: print(s, flush=True))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    peak, src = hbm_peak()
    json.dump({"peak_hbm_GBps": peak, "peak_source": src, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
