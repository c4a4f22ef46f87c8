"""oracle/dense_ref.py (folded fp16 weights, the checker of the CUDA dense path) against oracle/keras_ref.py (unfolded
parameters, Keras layer semantics, float64): agreement to within the fp16 rounding of the folded weights pins weights.fold()
(batch-norm folding, deconvolution layout, head concatenations) and the graph wiring of dense_ref, layout errors of any
kind would show as O(1) differences."""
import numpy as np
import pytest
import torch

from oracle import dense_ref, keras_ref


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("arch,size", [(50, 128), (101, 128)])
def test_backbone_fpn_rpn(pkg, arch, size):
    params = pkg.weights.synthetic(arch)
    ref = dense_ref.Ref(pkg.weights.fold(params, arch), arch, act_half=False, device="cpu")
    ker = keras_ref.KerasModel(params, arch)
    img = np.random.default_rng(1).integers(0, 256, (2, size, size, 3), dtype=np.uint8)
    with torch.no_grad():
        ps, probs, deltas = ref.backbone(img)
        kps, kprobs, kdeltas = ker.backbone(img)
    n = sum(3 * (size // s) ** 2 for s in (4, 8, 16, 32, 64))
    assert probs.shape == (2, n, 2) and tuple(kprobs.shape) == (2, n, 2) and tuple(kdeltas.shape) == (2, n, 4)
    for l in range(4):
        got = ps[l].permute(0, 3, 1, 2).numpy()                  # dense_ref returns NHWC
        assert got.shape == tuple(kps[l].shape)
        assert _rel(got, kps[l].numpy()) < 5e-3, f"P{l + 2}"           # measured 4e-4 .. 9e-4: the fp16 rounding of the folded weights
    assert np.abs(probs.numpy() - kprobs.numpy()).max() < 5e-3             # measured 8e-4 .. 1.4e-3
    assert _rel(deltas.numpy(), kdeltas.numpy()) < 5e-3


def test_classifier_and_mask_heads(pkg):
    params = pkg.weights.synthetic(50)
    ref = dense_ref.Ref(pkg.weights.fold(params, 50), 50, act_half=False, device="cpu")
    ker = keras_ref.KerasModel(params, 50)
    rng = np.random.default_rng(2)
    pooled7 = rng.standard_normal((6, 7, 7, 256)).astype(np.float16).astype(np.float32)
    pooled14 = rng.standard_normal((3, 14, 14, 256)).astype(np.float16).astype(np.float32)
    with torch.no_grad():
        p, bb, _ = ref.classifier(pooled7)
        kp, kbb = ker.classifier(pooled7)
        m = ref.mask(pooled14)
        km = ker.mask(pooled14)
    assert p.shape == (6, 81) and bb.shape == (6, 324) and m.shape == (3, 81, 28, 28)
    assert np.abs(p.numpy() - kp.numpy()).max() < 5e-3 and (p.argmax(1) == kp.argmax(1)).all()
    assert _rel(bb.numpy(), kbb.numpy()) < 1e-2
    assert np.abs(m.numpy() - km.numpy()).max() < 5e-3


def test_a_layout_error_would_be_seen(pkg):
    """Sensitivity check of the comparison itself: a transposed deconvolution kernel or swapped head halves are far outside
    the tolerances used above."""
    params = pkg.weights.synthetic(50)
    ker = keras_ref.KerasModel(params, 50)
    bad = {k: dict(v) for k, v in params.items()}
    bad["mask.deconv"]["kernel"] = np.ascontiguousarray(params["mask.deconv"]["kernel"][::-1])         # dy flipped
    ref_bad = dense_ref.Ref(pkg.weights.fold(bad, 50), 50, act_half=False, device="cpu")
    pooled14 = np.random.default_rng(3).standard_normal((2, 14, 14, 256)).astype(np.float32)
    with torch.no_grad():
        assert np.abs(ref_bad.mask(pooled14).numpy() - ker.mask(pooled14).numpy()).max() > 5e-2
