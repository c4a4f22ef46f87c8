#!/usr/bin/env python
"""Per-CTA role timeline of one fused expansion + reduction launch (csrc/conv_fused.cuh; debug).
usage: python tools/trace_fused.py n h w c1 n1 n2"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m

NAMES = {3: "M.a_landed", 10: "M.gemm1_start", 11: "M.gemm1_issued", 12: "M.gemm2_issued", 5: "E.acc1_ready", 9: "E.res_landed",
         6: "E.sub_staged", 15: "E.tmem_half", 16: "E.converted", 13: "E.acc2_ready", 14: "E.y_staged"}


def main():
    n, h, w, c1, n1, n2 = [int(x) for x in sys.argv[1:7]]
    ctx = m.Context()
    lib = m.lib()
    g = torch.Generator(device="cpu").manual_seed(1)
    a = torch.randn(n, h, w, c1, generator=g).half().cuda()
    w1 = (torch.randn(n1, c1, generator=g) / np.sqrt(c1)).half().cuda()
    b1 = torch.randn(n1, generator=g).cuda()
    res = torch.randn(n, h, w, n1, generator=g).half().cuda()
    w2 = (torch.randn(n2, n1, generator=g) / np.sqrt(n1)).half().cuda()
    b2 = torch.randn(n2, generator=g).cuda()
    x = torch.empty((n, h, w, n1), dtype=torch.float16, device="cuda")
    y = torch.empty((n, h, w, n2), dtype=torch.float16, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    SEC = 2 * 340 + 2
    trace = torch.zeros(148 * 3 * SEC, dtype=torch.int64, device="cuda")
    for it in range(3):
        flush.zero_()
        trace.zero_()
        torch.cuda.synchronize()
        lib.mrcnn_debug_conv_trace(trace.data_ptr() if it == 2 else None)
        m._cabi.check(ctx.handle, lib.mrcnn_debug_fused_expand_reduce(ctx.handle, a.data_ptr(), n, h, w, c1, w1.data_ptr(), b1.data_ptr(), n1,
                                                                      res.data_ptr(), w2.data_ptr(), b2.data_ptr(), n2, x.data_ptr(), y.data_ptr()))
        ctx.synchronize()
    lib.mrcnn_debug_conv_trace(None)
    t = trace.cpu().numpy().reshape(148, 3, SEC)
    for cta in (0, 73):
        ev = []
        for role in range(3):
            cnt = int(t[cta, role, 0])
            ev += [(int(t[cta, role, 2 + 2 * i]) >> 32, int(t[cta, role, 2 + 2 * i]) & 0xffffffff, int(t[cta, role, 3 + 2 * i])) for i in range(min(cnt, 340))]
        if not ev:
            continue
        ev.sort(key=lambda e: e[2])
        t0 = ev[0][2]
        print(f"--- CTA {cta}: {len(ev)} events, span {ev[-1][2] - t0} clk")
        last = {}
        for code, arg, clk in ev:
            d = clk - last.get(code, t0)
            last[code] = clk
            print(f"  {clk - t0:8d}  (+{d:6d})  {NAMES.get(code, code)} {arg}")


if __name__ == "__main__":
    main()
