#!/usr/bin/env python
"""Per-CTA role timeline of one convolution launch (debug): producer / MMA / epilogue events with clock64 stamps.
usage: python tools/trace_conv.py n h w cin cout k stride [ctas] [residual: 0|1]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m

NAMES = {0: "prologue_done", 7: "dep_ok", 1: "P.issue", 2: "M.landed", 3: "M.tile_start", 4: "M.tile_issued", 5: "E.acc_ready", 9: "E.buf_ready", 6: "E.chunk_staged", 8: "E.done"}


def main():
    n, h, w, cin, cout, k, stride = [int(x) for x in sys.argv[1:8]]
    if len(sys.argv) > 8:
        if int(sys.argv[8]):
            os.environ["MRCNN_CONV_CTAS"] = sys.argv[8]
    ctx = m.Context()
    st = torch.cuda.Stream()
    ctx.set_stream(st.cuda_stream)
    lib = m.lib()
    pad = k // 2
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    x = torch.randn((n, h, w, cin), device="cuda", dtype=torch.float16)
    wt = torch.randn((cout, k, k, cin), device="cuda", dtype=torch.float16) * 0.05
    bias = torch.zeros(cout, device="cuda")
    out = torch.empty((n, ho, wo, (cout + 7) // 8 * 8), device="cuda", dtype=torch.float16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = None
    if len(sys.argv) > 9 and int(sys.argv[9]):
        res = torch.randn((n, ho, wo, cout), device="cuda", dtype=torch.float16)
    SEC = 2 * 340 + 2
    trace = torch.zeros(148 * 3 * SEC, dtype=torch.int64, device="cuda")
    with torch.cuda.stream(st):
        for it in range(3):
            flush.zero_()
            trace.zero_()
            st.synchronize()
            lib.mrcnn_debug_conv_trace(trace.data_ptr() if it == 2 else None)
            rc = lib.mrcnn_conv2d_nhwc_f16(ctx.handle, x.data_ptr(), n, h, w, cin, wt.data_ptr(), bias.data_ptr(), cout, k, k, stride, pad, res.data_ptr() if res is not None else None, 1, out.data_ptr())
            m._cabi.check(ctx.handle, rc)
            st.synchronize()
    lib.mrcnn_debug_conv_trace(None)
    t = trace.cpu().numpy().reshape(148, 3, SEC)
    for cta in (0, 1, 73):
        ev = []
        for role in range(3):
            cnt = int(t[cta, role, 0])
            ev += [(int(t[cta, role, 2 + 2 * i]) >> 32, int(t[cta, role, 2 + 2 * i]) & 0xffffffff, int(t[cta, role, 3 + 2 * i])) for i in range(min(cnt, 340))]
        cnt = len(ev)
        if not ev:
            continue
        ev.sort(key=lambda e: e[2])
        t0 = ev[0][2]
        print(f"--- CTA {cta}: {cnt} events, span {ev[-1][2] - t0} clk")
        last = {}
        for code, arg, clk in ev:
            d = clk - last.get(code, t0)
            last[code] = clk
            print(f"  {clk - t0:8d}  (+{d:6d} since last {NAMES.get(code, code)})  {NAMES.get(code, code)} {arg}")


if __name__ == "__main__":
    main()
