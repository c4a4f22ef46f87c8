"""GPU numerics of the tcgen05 implicit-GEMM convolution against a plain PyTorch
fp32 reference of the same op (fp16 inputs, fp32 accumulate)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _conv_ref(x, w, bias, stride, pad, residual, relu):
    import torch
    import torch.nn.functional as F
    # a real fp32 reference: cuDNN / cuBLAS may otherwise run fp32 convolutions in TF32 (10-bit mantissa, the precision
    # of the kernel under test) on this GPU
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, stride=stride, padding=pad)
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    if relu:
        y = torch.relu(y)
    return y


CASES = [
    # n, h, w, cin, cout, k, stride, pad, residual, relu
    (2, 32, 32, 64, 64, 1, 1, 0, False, True),
    (1, 32, 48, 128, 256, 3, 1, 1, True, True),
    (2, 16, 16, 256, 512, 1, 2, 0, False, False),
    (1, 14, 14, 256, 256, 3, 1, 1, False, True),       # mask-head shape: partial tiles in x and y
    (1, 8, 8, 64, 24, 1, 1, 0, False, False),          # tiny map, cout not a multiple of 32
    (1, 64, 64, 64, 128, 3, 1, 1, False, True),
    (1, 1, 300, 1024, 408, 1, 1, 0, False, False),     # GEMM-like (FC layers): h = 1
    # staged (TMA store) epilogue: many tiles per CTA, 1..4 chunks of 64 channels, TMA-prefetched residual
    (8, 64, 64, 64, 256, 1, 1, 0, True, True),
    (4, 40, 56, 128, 512, 1, 1, 0, True, True),        # two N tiles x four chunks, partial tiles
    (2, 30, 30, 256, 128, 3, 1, 1, True, False),       # BN = 128, ragged edges
    (3, 64, 64, 256, 512, 1, 2, 0, False, True),       # stride 2 (first 1x1 of a stage)
    (2, 14, 14, 64, 64, 3, 1, 1, True, True),          # single chunk with residual
    (1, 1, 1000, 12544, 1024, 1, 1, 0, False, True),   # classifier-head GEMM (long K)
]


@pytest.mark.parametrize("case", CASES)
def test_conv_matches_torch(pkg, ctx, case):
    import torch
    n, h, w, cin, cout, k, stride, pad, use_res, relu = case
    g = torch.Generator(device="cpu").manual_seed(hash(case) & 0xffff)
    x = (torch.randn(n, h, w, cin, generator=g) * 1.0).half().cuda()
    wt = (torch.randn(cout, k, k, cin, generator=g) / np.sqrt(k * k * cin)).half().cuda()
    bias = torch.randn(cout, generator=g).cuda()
    ho = (h + 2 * pad - k) // stride + 1
    wo = (w + 2 * pad - k) // stride + 1
    ldc = (cout + 7) // 8 * 8
    res = (torch.randn(n, ho, wo, ldc, generator=g)).half().cuda() if use_res else None
    out = torch.full((n, ho, wo, ldc), 77.0, dtype=torch.float16, device="cuda")
    rc = pkg.lib().mrcnn_conv2d_nhwc_f16(ctx.handle, x.data_ptr(), n, h, w, cin, wt.data_ptr(), bias.data_ptr(), cout, k, k,
                                         stride, pad, res.data_ptr() if use_res else None, int(relu), out.data_ptr())
    pkg._cabi.check(ctx.handle, rc)
    ctx.synchronize()
    ref = _conv_ref(x, wt, bias, stride, pad, res[..., :cout] if use_res else None, relu)
    got = out[..., :cout].float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * max(scale, 1.0), f"max abs err {err} (scale {scale})"


@pytest.mark.parametrize("relu", [True, False])
def test_conv_epilogue_rounds_once(pkg, ctx, relu):
    """The epilogue contract: out = fp16(relu(acc + bias + residual)) with ONE rounding (fp32 sum, packed / mixed-precision
    adds in the kernel).  Against an fp64 evaluation rounded once to fp16: at most one fp16 ulp away (the fp32 accumulation
    order differs), almost always equal; clamped outputs are +0, never -0."""
    import torch
    n, h, w, cin, cout = 4, 32, 32, 128, 256
    g = torch.Generator(device="cpu").manual_seed(1234 + relu)
    x = torch.randn(n, h, w, cin, generator=g).half().cuda()
    wt = (torch.randn(cout, 1, 1, cin, generator=g) / np.sqrt(cin)).half().cuda()
    bias = torch.randn(cout, generator=g).cuda()
    res = torch.randn(n, h, w, cout, generator=g).half().cuda()
    out = torch.full((n, h, w, cout), 77.0, dtype=torch.float16, device="cuda")
    rc = pkg.lib().mrcnn_conv2d_nhwc_f16(ctx.handle, x.data_ptr(), n, h, w, cin, wt.data_ptr(), bias.data_ptr(), cout, 1, 1,
                                         1, 0, res.data_ptr(), int(relu), out.data_ptr())
    pkg._cabi.check(ctx.handle, rc)
    ctx.synchronize()
    ref = x.double().reshape(-1, cin) @ wt.double().reshape(cout, cin).t() + bias.double() + res.double().reshape(-1, cout)
    if relu:
        ref = torch.relu(ref)
    want = ref.half().reshape(n, h, w, cout)
    gb = out.view(torch.int16).int()
    wb = want.view(torch.int16).int()
    # fp16 bit patterns of equal sign are ordered like the values: a difference of 1 in the pattern is one ulp
    same_sign = (gb < 0) == (wb < 0)
    ulp = (gb - wb).abs()
    near_zero = (out.float().abs() < 1e-3) & (want.float().abs() < 1e-3)       # sign may flip across zero
    assert bool(((ulp <= 1) & same_sign | near_zero).all()), f"max ulp distance {int(ulp[same_sign].max())}"
    assert float((gb == wb).float().mean()) > 0.995
    if relu:
        assert int((gb < 0).sum()) == 0, "ReLU outputs must not carry a sign bit (-0)"
        assert bool((out[ref.reshape(n, h, w, cout) <= 0] == 0).all())


def test_conv_grid_variants_bit_identical(tmp_path):
    """MRCNN_CONV_BALANCED (balanced persistent grid vs one CTA per SM) and MRCNN_CONV_CTAS (CTA pairs vs single CTAs) change
    the schedule, never the result (same K order per output row): the knobs are read once per process, so each variant runs
    in its own interpreter.  (MRCNN_CONV_VGROUP reorders the taps, i.e. the fp32 summation: equal only within tolerance.)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = r'''
import sys, zlib
sys.path.insert(0, %r)
import numpy as np, torch
import maskrcnn_b200 as m
ctx = m.Context()
crc = 0
for (n, h, w, cin, cout, k, res) in [(8, 64, 64, 256, 1024, 1, True), (8, 64, 64, 256, 256, 3, False), (3, 50, 38, 64, 128, 3, True)]:
    g = torch.Generator(device="cpu").manual_seed(7)
    x = torch.randn(n, h, w, cin, generator=g).half().cuda()
    wt = (torch.randn(cout, k, k, cin, generator=g) / np.sqrt(k * k * cin)).half().cuda()
    bias = torch.randn(cout, generator=g).cuda()
    r = torch.randn(n, h, w, cout, generator=g).half().cuda() if res else None
    out = torch.zeros((n, h, w, cout), dtype=torch.float16, device="cuda")
    rc = m.lib().mrcnn_conv2d_nhwc_f16(ctx.handle, x.data_ptr(), n, h, w, cin, wt.data_ptr(), bias.data_ptr(), cout, k, k, 1, k // 2,
                                       r.data_ptr() if res else None, 1, out.data_ptr())
    m._cabi.check(ctx.handle, rc)
    ctx.synchronize()
    crc = zlib.crc32(out.cpu().numpy().tobytes(), crc)
print("CRC", crc)
''' % root
    crcs = {}
    for name, env in [("default", {}), ("one CTA per SM", {"MRCNN_CONV_BALANCED": "0"}), ("single CTAs", {"MRCNN_CONV_CTAS": "1"})]:
        e = dict(os.environ)
        e.update(env)
        p = subprocess.run([sys.executable, "-c", script], env=e, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        crcs[name] = [ln for ln in p.stdout.splitlines() if ln.startswith("CRC")][-1]
    assert len(set(crcs.values())) == 1, crcs


FUSED_CASES = [
    # n, h, w, c1, n1, n2
    (8, 64, 64, 256, 1024, 256),       # res4 at batch 8: 256 tiles, two per CTA
    (2, 30, 38, 128, 512, 128),        # res3-like, ragged tiles
    (1, 16, 16, 64, 256, 64),          # res2-like, two chunks per tile, a single Y sub-chunk
    (3, 64, 64, 256, 1024, 256),
    (4, 64, 128, 64, 256, 64),         # several tiles per CTA with one Y sub-chunk (the store thread frees it at the tile's end)
    (8, 64, 96, 128, 512, 128),        # several tiles per CTA, two Y sub-chunks
    (8, 128, 128, 256, 1024, 256),     # seven tiles per CTA
    (1, 24, 16, 128, 512, 128),        # three tiles: the second CTA pair owns a phantom tile
]


@pytest.mark.parametrize("case", FUSED_CASES)
def test_fused_expand_reduce_bit_identical(pkg, ctx, case):
    """csrc/conv_fused.cuh: the 1x1 expansion (+ residual) of a bottleneck block and the 1x1 reduction of the next block in
    one launch == the two separate launches, bit for bit (X and Y)."""
    import torch
    n, h, w, c1, n1, n2 = case
    g = torch.Generator(device="cpu").manual_seed(hash(case) & 0xffff)
    a = torch.randn(n, h, w, c1, generator=g).half().cuda()
    w1 = (torch.randn(n1, 1, 1, c1, generator=g) / np.sqrt(c1)).half().cuda()
    b1 = torch.randn(n1, generator=g).cuda()
    res = torch.randn(n, h, w, n1, generator=g).half().cuda()
    w2 = (torch.randn(n2, 1, 1, n1, generator=g) / np.sqrt(n1)).half().cuda()
    b2 = torch.randn(n2, generator=g).cuda()
    lib = pkg.lib()
    x_ref = torch.full((n, h, w, n1), 7.0, dtype=torch.float16, device="cuda")
    y_ref = torch.full((n, h, w, n2), 7.0, dtype=torch.float16, device="cuda")
    torch.cuda.synchronize()
    pkg._cabi.check(ctx.handle, lib.mrcnn_conv2d_nhwc_f16(ctx.handle, a.data_ptr(), n, h, w, c1, w1.data_ptr(), b1.data_ptr(), n1, 1, 1, 1, 0,
                                                          res.data_ptr(), 1, x_ref.data_ptr()))
    pkg._cabi.check(ctx.handle, lib.mrcnn_conv2d_nhwc_f16(ctx.handle, x_ref.data_ptr(), n, h, w, n1, w2.data_ptr(), b2.data_ptr(), n2, 1, 1, 1, 0,
                                                          None, 1, y_ref.data_ptr()))
    x = torch.full((n, h, w, n1), 9.0, dtype=torch.float16, device="cuda")
    y = torch.full((n, h, w, n2), 9.0, dtype=torch.float16, device="cuda")
    torch.cuda.synchronize()           # the context's stream is non-blocking: the fills above must not race the kernel
    pkg._cabi.check(ctx.handle, lib.mrcnn_debug_fused_expand_reduce(ctx.handle, a.data_ptr(), n, h, w, c1, w1.data_ptr(), b1.data_ptr(), n1,
                                                                    res.data_ptr(), w2.data_ptr(), b2.data_ptr(), n2, x.data_ptr(), y.data_ptr()))
    ctx.synchronize()
    assert torch.equal(x, x_ref), f"X differs in {(x != x_ref).sum().item()} of {x.numel()} elements"
    assert torch.equal(y, y_ref), f"Y differs in {(y != y_ref).sum().item()} of {y.numel()} elements"


def test_fused_expand_reduce_cta_pair_variant():
    """MRCNN_FUSE_CTAS=2 (the fused kernel on CTA pairs, tcgen05.mma.cta_group::2; off by default): same bits as the
    single-CTA variant.  The knob is read once per process, so the variant runs in its own interpreter."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = r'''
import sys, zlib
sys.path.insert(0, %r)
import numpy as np, torch
import maskrcnn_b200 as m
ctx = m.Context()
crc = 0
for (n, h, w, c1, n1, n2) in [(8, 64, 64, 256, 1024, 256), (1, 24, 16, 128, 512, 128), (4, 64, 128, 64, 256, 64)]:
    g = torch.Generator(device="cpu").manual_seed(11)
    a = torch.randn(n, h, w, c1, generator=g).half().cuda()
    w1 = (torch.randn(n1, c1, generator=g) / np.sqrt(c1)).half().cuda(); b1 = torch.randn(n1, generator=g).cuda()
    res = torch.randn(n, h, w, n1, generator=g).half().cuda()
    w2 = (torch.randn(n2, n1, generator=g) / np.sqrt(n1)).half().cuda(); b2 = torch.randn(n2, generator=g).cuda()
    x = torch.zeros((n, h, w, n1), dtype=torch.float16, device="cuda"); y = torch.zeros((n, h, w, n2), dtype=torch.float16, device="cuda")
    torch.cuda.synchronize()
    m._cabi.check(ctx.handle, m.lib().mrcnn_debug_fused_expand_reduce(ctx.handle, a.data_ptr(), n, h, w, c1, w1.data_ptr(), b1.data_ptr(), n1,
                                                                        res.data_ptr(), w2.data_ptr(), b2.data_ptr(), n2, x.data_ptr(), y.data_ptr()))
    ctx.synchronize()
    crc = zlib.crc32(y.cpu().numpy().tobytes(), zlib.crc32(x.cpu().numpy().tobytes(), crc))
print("CRC", crc)
''' % root
    crcs = {}
    for ctas in ("1", "2"):
        e = dict(os.environ)
        e["MRCNN_FUSE_CTAS"] = ctas
        p = subprocess.run([sys.executable, "-c", script], env=e, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        crcs[ctas] = [ln for ln in p.stdout.splitlines() if ln.startswith("CRC")][-1]
    assert crcs["1"] == crcs["2"], crcs
