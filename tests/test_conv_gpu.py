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
