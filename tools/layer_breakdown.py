#!/usr/bin/env python
"""Joins an ncu launch list of ONE warm batch-8 step (profiles/*_pipeline_launches_warm.csv: duration, DRAM bytes and
tensor-pipe activity per launch) with the layer plan of the dense graphs (execution order of csrc/pipeline.cu), and prints
per layer class: launches, time, share of the step, achieved TFLOP/s, tensor-pipe activity, tiles per launch and the
wave count on 148 SMs, and the time the class would take at the measured sustained peak (MEASURED_PEAKS.json) -- i.e.
where the step's time goes and how much of it is above the roofline.  Pure post-processing: no GPU needed.
  python tools/layer_breakdown.py profiles/r1w_pipeline_launches_warm.csv [--traffic-json profiles/traffic_r2.json] > profiles/r1w_conv_layer_breakdown.txt"""
import collections
import csv
import io
import json
import math
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, IMG, R, D = 8, 1024, 1000, 100


ONE_TERM_MASKS = "--one-term-masks" in sys.argv      # launch lists taken with precise_masks = 0 (round 1)
UNFUSED = "--unfused" in sys.argv                    # launch lists taken before csrc/conv_fused.cuh (or with MRCNN_CONV_FUSE=0)


def plan(architecture=101):
    """[(class, M, N, K_useful, algorithmic bytes)] of every conv_gemm launch of one step, in launch order.
    Algorithmic bytes = every operand once: input pixels the layer needs (fp16), weights, output, residual."""
    L = []

    def add(cls, M, N, K, cin, extra=0, out_bytes=None):
        L.append((cls, M, N, K, 2 * M * cin + 2 * N * K + (2 * M * N if out_bytes is None else out_bytes) + extra, 2.0 * M * N * K))
    h = IMG // 2
    add("stem 7x7/2", B * h * h, 64, 147, 16)                       # space-to-depth image: 16 fp16 per output pixel
    h //= 2
    cin = 64
    for s, nb in enumerate({101: (3, 4, 23, 3), 50: (3, 4, 6, 3)}[architecture]):
        f = 64 << s
        fusable = (not UNFUSED) and f <= 256            # csrc/conv_fused.cuh: A tile of the expansion <= 256 channels
        for i in range(nb):
            stride = 2 if (i == 0 and s > 0) else 1
            ho = h // stride
            m = B * ho * ho
            if not (fusable and i > 0):                   # else: this block's reduction ran inside the previous block's fused launch
                add(f"res{s + 2} 2a 1x1", m, f, cin, cin)
            add(f"res{s + 2} 2b 3x3", m, f, 9 * f, f)
            if i == 0:
                add(f"res{s + 2} shortcut 1x1", m, 4 * f, cin, cin)
            if fusable and i < nb - 1:
                # expansion (+ residual) and the next block's reduction in one launch: X is written once and not read back
                L.append((f"res{s + 2} 2c + next 2a (fused)", m, 4 * f, f,
                          2 * m * f + 2 * m * 4 * f + 2 * m * 4 * f + 2 * m * f + 2 * 2 * (4 * f * f), 2.0 * m * (4 * f * f + f * 4 * f)))
            else:
                add(f"res{s + 2} 2c 1x1 + residual", m, 4 * f, f, f, extra=2 * m * 4 * f)
            h, cin = ho, 4 * f
    lv = [IMG // 4 >> l for l in range(5)]
    for l, c in zip((3, 2, 1, 0), (2048, 1024, 512, 256)):
        m = B * lv[l] ** 2
        add("fpn lateral 1x1 (+ top-down)", m, 256, c, c, extra=0 if l == 3 else 2 * (m // 4) * 256)
    for l in range(4):
        add("fpn output 3x3", B * lv[l] ** 2, 256, 9 * 256, 256)
    for l in range(5):
        m = B * lv[l] ** 2
        add("rpn shared 3x3", m, 512, 9 * 256, 256)
        add("rpn head 1x1 (18 ch)", m, 18, 512, 512, out_bytes=4 * m * 24)          # fp32 rows of 24
    add("classifier conv1 7x7 (gemm)", B * R, 1024, 49 * 256, 49 * 256)
    add("classifier conv2 1x1", B * R, 1024, 1024, 1024)
    add("classifier logits+boxes", B * R, 405, 1024, 1024, out_bytes=4 * B * R * 408)
    m = B * D * 14 * 14
    if ONE_TERM_MASKS:
        for _ in range(4):
            add("mask conv 3x3", m, 256, 9 * 256, 256)
        add("mask deconv 2x2 + class plane", m, 1024, 256, 256, out_bytes=4 * m * 4)     # only the selected class plane leaves, fp32
    else:
        # library default (precise_masks = 1): activations leave every layer as (hi, lo) fp16 pairs = 512 channels, and the
        # layers that consume them run with K doubled (hi * w + lo * w); the pooled input of conv1 is exact fp16 (K = 9 * 256)
        add("mask conv 3x3", m, 256, 9 * 256, 256, out_bytes=2 * m * 512)
        for _ in range(3):
            add("mask conv 3x3", m, 256, 9 * 512, 512, out_bytes=2 * m * 512)
        add("mask deconv 2x2 + class plane", m, 1024, 512, 512, out_bytes=4 * m * 4)
    return L


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    path = args[0] if args else os.path.join(ROOT, "profiles", "r1w_pipeline_launches_warm.csv")
    txt = open(path).read()
    txt = txt[txt.index('"ID"'):]
    by = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO(txt)):
        d = by.setdefault(int(row["ID"]), {"name": row["Kernel Name"]})
        d[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    launches = list(by.values())
    start = next(i for i, d in enumerate(launches) if d["name"].startswith("preprocess_s2d"))
    launches = launches[start:] + launches[:start]          # the capture may start mid-step; a step begins at the pre-processing
    convs = [d for d in launches if "conv_gemm_kernel" in d["name"] or "conv_fused_expand_reduce_kernel" in d["name"]]
    layers = plan()
    assert len(convs) == len(layers), (len(convs), len(layers))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1380.1))
    hbm = float(peaks.get("hbm_gbs", 6462.7))
    step_ns = sum(d["gpu__time_duration.sum"] for d in launches)
    agg = collections.OrderedDict()
    for d, (cls, M, N, K, alg_bytes, fl) in zip(convs, layers):
        mt = re.search(r"conv_gemm_kernel<(\d+), (\d+)>", d["name"])
        fused = mt is None
        assert fused == ("fused" in cls), (cls, d["name"])
        bn, ctas = (N, 1) if fused else map(int, mt.groups())        # the fused kernel: one 128-pixel tile, all channels
        tiles = math.ceil(M / (128 * ctas)) * math.ceil(N / bn)
        units = 148 // ctas                                  # CTA pairs: 74 clusters
        a = agg.setdefault(cls, dict(n=0, ns=0.0, flops=0.0, tens=0.0, dram=0.0, alg=0.0, bound=0.0, membound=0, tiles=tiles, waves=tiles / units,
                                     cfg="128xall" if fused else f"{128 * ctas}x{bn}"))
        a["n"] += 1
        a["ns"] += d["gpu__time_duration.sum"]
        a["flops"] += fl
        a["tens"] += d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"] * d["gpu__time_duration.sum"]
        dram = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        a["dram"] += dram
        a["alg"] += alg_bytes
        # roofline of this launch: tensor time at the sustained peak vs the bytes it cannot avoid moving (L2-resident
        # operands make the measured DRAM bytes smaller than the algorithmic ones; wasted re-reads make them larger)
        t_c, t_m = fl / (peak * 1e3), min(alg_bytes, dram) / hbm
        a["bound"] += max(t_c, t_m)
        a["membound"] += t_m > t_c
    conv_ns = sum(a["ns"] for a in agg.values())
    print(f"# {os.path.basename(path)}: one warm batch-{B} step under ncu (per-launch times are serialised and cold-ish: use SHARES, not absolutes)")
    print(f"# step {step_ns / 1e6:.3f} ms in {len(launches)} launches; conv_gemm_kernel {conv_ns / 1e6:.3f} ms ({100 * conv_ns / step_ns:.1f} %) in {len(convs)} launches; "
          f"peak = {peak:.1f} TFLOP/s (sustained, MEASURED_PEAKS.json)")
    print(f"# roofline ms = sum over launches of max(flops / {peak:.0f} TFLOP/s, min(algorithmic, measured DRAM) bytes / {hbm:.0f} GB/s); 'hbm' = launches bound by the second term")
    print(f"{'layer class':34s} {'n':>3s} {'tile':>8s} {'tiles':>6s} {'waves':>6s} {'ms':>7s} {'share':>6s} {'TFLOP/s':>8s} {'tensor%':>7s} {'DRAM MB':>8s} {'alg MB':>8s} {'GB/s':>6s} {'hbm':>5s} {'roof ms':>8s} {'frac':>5s} {'excess ms':>9s}")
    rows = sorted(agg.items(), key=lambda kv: -kv[1]["ns"])
    tot_excess = tot_roof = 0.0
    for cls, a in rows:
        ms = a["ns"] / 1e6
        tf = a["flops"] / a["ns"] / 1e3
        roof = a["bound"] / 1e6
        tot_excess += ms - roof
        tot_roof += roof
        print(f"{cls:34s} {a['n']:3d} {a['cfg']:>8s} {a['tiles']:6d} {a['waves']:6.2f} {ms:7.3f} {100 * a['ns'] / step_ns:5.1f}% {tf:8.1f} "
              f"{a['tens'] / a['ns']:7.1f} {a['dram'] / a['n'] / 1e6:8.1f} {a['alg'] / a['n'] / 1e6:8.1f} {a['dram'] / a['ns']:6.0f} {a['membound']:2d}/{a['n']:<2d} {roof:8.3f} {roof / ms:5.2f} {ms - roof:9.3f}")
    flops = sum(a["flops"] for a in agg.values())
    print(f"{'all conv launches':34s} {len(convs):3d} {'':>8s} {'':>6s} {'':>6s} {conv_ns / 1e6:7.3f} {100 * conv_ns / step_ns:5.1f}% {flops / conv_ns / 1e3:8.1f} "
          f"{sum(a['tens'] for a in agg.values()) / conv_ns:7.1f} {'':>8s} {'':>8s} {sum(a['dram'] for a in agg.values()) / conv_ns:6.0f} {'':>5s} {tot_roof:8.3f} {tot_roof / (conv_ns / 1e6):5.2f} {tot_excess:9.3f}")
    print(f"# pure tensor roofline (all flops at {peak:.0f} TFLOP/s): {flops / (peak * 1e12) * 1e3:.3f} ms = {flops / conv_ns / 1e3 / peak:.2f} of the measured conv time")
    other = collections.Counter()
    for d in launches:
        if "conv_gemm_kernel" not in d["name"]:
            other[d["name"].split("(")[0].replace("void ", "")] += d["gpu__time_duration.sum"]
    print("\n# other kernels (ms per step)")
    for k, v in other.most_common():
        print(f"{k:34s} {v / 1e6:7.3f} {100 * v / step_ns:5.1f}%")
    if "--traffic-json" in sys.argv:
        # per-class DRAM traffic of the step (bench.py reads roofline.traffic from this file)
        out = sys.argv[sys.argv.index("--traffic-json") + 1]
        rd = sum(d["dram__bytes_read.sum"] for d in convs)
        wr = sum(d["dram__bytes_write.sum"] for d in convs)
        ra = [d for d in launches if "roialign_staged_kernel" in d["name"] or "roialign_nhwc_kernel" in d["name"]]
        json.dump({
            "source": f"profiles/{os.path.basename(path)} ({len(launches)} launches = one warm batch-{B} step): ncu --metrics gpu__time_duration.sum,"
                      "dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active... --clock-control none --cache-control none "
                      "(python bench.py --steps 1 --warmup 3); written by tools/layer_breakdown.py --traffic-json",
            "conv_gemm_tcgen05": {
                "launches_per_step": len(convs), "dram_bytes_read_per_step": int(rd), "dram_bytes_write_per_step": int(wr),
                "traffic_bytes_per_launch": int((rd + wr) / len(convs)),
                "note": "average over the conv launches of a step (layer sizes differ by 3 orders of magnitude)",
                "share_of_step_time_ncu": round(conv_ns / step_ns, 3),
                "tensor_pipe_active_time_weighted_pct": round(sum(a["tens"] for a in agg.values()) / conv_ns, 1)},
            "roialign": {
                "launches_per_step": len(ra),
                "traffic_bytes_per_launch": int(sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in ra) / max(len(ra), 1)),
                "per_launch": [{"kernel": d["name"].split("(")[0], "dram_read": d["dram__bytes_read.sum"], "dram_write": d["dram__bytes_write.sum"],
                                "us_under_ncu": d["gpu__time_duration.sum"] / 1e3} for d in ra]},
        }, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
