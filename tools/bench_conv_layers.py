#!/usr/bin/env python
"""Per-layer microbench of the tcgen05 implicit-GEMM convolution on the distinct layer shapes of the
ResNet101+FPN+RPN / head graphs at batch 8, 1024x1024 (CUDA-event timed through mrcnn_conv2d_nhwc_f16).
Prints one line per distinct shape: count per step, ms, TFLOP/s, effective GB/s, and the share of the step."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m


def layer_shapes(batch=8, size=1024, arch=101):
    """[(tag, n, h, w, cin, cout, k, stride, count)]"""
    out = []
    nb = {101: (3, 4, 23, 3), 50: (3, 4, 6, 3)}[arch]
    h = size // 4
    cin = 64
    for s, n in enumerate(nb):
        f = 64 << s
        for i in range(n):
            st = 2 if (i == 0 and s > 0) else 1
            ho = h // st
            out.append((f"res{s+2} 2a", batch, h, h, cin, f, 1, st))
            out.append((f"res{s+2} 2b", batch, ho, ho, f, f, 3, 1))
            if i == 0:
                out.append((f"res{s+2} sc", batch, h, h, cin, 4 * f, 1, st))
            out.append((f"res{s+2} 2c", batch, ho, ho, f, 4 * f, 1, 1))
            h, cin = ho, 4 * f
    for l, c in zip((2, 3, 4, 5), (256, 512, 1024, 2048)):
        hh = size // (2 ** l)
        out.append((f"fpn lat{l}", batch, hh, hh, c, 256, 1, 1))
        out.append((f"fpn p{l}", batch, hh, hh, 256, 256, 3, 1))
    for l in (2, 3, 4, 5, 6):
        hh = size // (2 ** l)
        out.append((f"rpn shared p{l}", batch, hh, hh, 256, 512, 3, 1))
        out.append((f"rpn head p{l}", batch, hh, hh, 512, 18, 1, 1))
    out.append(("cls conv1 (gemm)", 1, 1, batch * 1000, 12544, 1024, 1, 1))
    out.append(("cls conv2 (gemm)", 1, 1, batch * 1000, 1024, 1024, 1, 1))
    out.append(("cls fc (gemm)", 1, 1, batch * 1000, 1024, 405, 1, 1))
    for i in range(4):
        out.append(("mask conv3x3", batch * 100, 14, 14, 256, 256, 3, 1))
    agg = {}
    for t in out:
        key = t[1:]
        if key in agg:
            agg[key][1] += 1
        else:
            agg[key] = [t[0], 1]
    return [(v[0],) + k + (v[1],) for k, v in agg.items()]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="", help="comma separated substrings of layer tags to run")
    args = ap.parse_args()
    ctx = m.Context()
    st = torch.cuda.Stream()
    ctx.set_stream(st.cuda_stream)
    lib = m.lib()
    rows = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for tag, n, h, w, cin, cout, k, stride, count in layer_shapes(args.batch):
        if args.only and not any(o in tag for o in args.only.split(",")):
            continue
        pad = k // 2
        ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        ldc = (cout + 7) // 8 * 8
        x = torch.randn((n, h, w, cin), device="cuda", dtype=torch.float16)
        wt = torch.randn((cout, k, k, cin), device="cuda", dtype=torch.float16) * 0.05
        bias = torch.zeros(cout, device="cuda")
        out = torch.empty((n, ho, wo, ldc), device="cuda", dtype=torch.float16)
        # the last 1x1 of every bottleneck adds the block input (TMA-fetched residual epilogue)
        res = torch.randn((n, ho, wo, ldc), device="cuda", dtype=torch.float16) if tag.endswith(" 2c") else None
        ts = []
        with torch.cuda.stream(st):
            for r in range(args.reps + 2):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.mrcnn_conv2d_nhwc_f16(ctx.handle, x.data_ptr(), n, h, w, cin, wt.data_ptr(), bias.data_ptr(), cout, k, k, stride,
                                               pad, res.data_ptr() if res is not None else None, 1, out.data_ptr())
                e1.record()
                m._cabi.check(ctx.handle, rc)
                st.synchronize()
                if r >= 2:
                    ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        flops = 2.0 * n * ho * wo * cout * k * k * cin
        byts = 2.0 * (x.numel() + out.numel() + wt.numel() + (res.numel() if res is not None else 0))
        rows.append((tag, n, h, w, cin, cout, k, stride, count, ms, flops, byts))
        del x, wt, out, res
    total = sum(r[8] * r[9] for r in rows)
    print(f"{'layer':20s} {'n':>5s} {'hxw':>11s} {'cin':>6s} {'cout':>5s} k s  cnt    ms/launch  TFLOP/s   GB/s  ms/step share")
    for tag, n, h, w, cin, cout, k, stride, count, ms, flops, byts in sorted(rows, key=lambda r: -r[8] * r[9]):
        print(f"{tag:20s} {n:5d} {h:5d}x{w:<5d} {cin:6d} {cout:5d} {k} {stride} {count:4d} {ms:12.4f} {flops/ms/1e9:8.1f} {byts/ms/1e6:7.0f} {count*ms:8.3f} {100*count*ms/total:5.1f}%")
    print(f"sum over the step (cold L2 each launch): {total:.3f} ms, {sum(r[8]*r[10] for r in rows)/total/1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
