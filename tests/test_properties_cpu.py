"""Randomised (hypothesis) agreement of the two CPU restatements on the primitives every layer is built from: greedy NMS
with Double IoU / Float compare, box decoding op by op, the level rule, crop-and-resize, plus the invariants the CUDA
path is tested for at full size (tests/test_fullsize_gpu.py): idempotence of NMS, outputs inside [0, 1], zero padding."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import literal as lit
from oracle import oracle as orc

SET = settings(max_examples=60, deadline=None, derandomize=True)
coord = st.floats(0.0, 1.0, width=32)


@st.composite
def boxes(draw, max_n=40):
    n = draw(st.integers(0, max_n))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    kind = draw(st.sampled_from(["uniform", "clustered", "grid"]))
    if kind == "uniform":
        c = rng.uniform(0, 1, (n, 4)).astype(np.float32)
        y1, y2 = np.minimum(c[:, 0], c[:, 2]), np.maximum(c[:, 0], c[:, 2])
        x1, x2 = np.minimum(c[:, 1], c[:, 3]), np.maximum(c[:, 1], c[:, 3])
    elif kind == "clustered":                           # many overlaps close to the thresholds
        base = rng.uniform(0.2, 0.6, 4).astype(np.float32)
        j = rng.uniform(-0.05, 0.05, (n, 4)).astype(np.float32)
        y1, x1 = base[0] + j[:, 0], base[1] + j[:, 1]
        y2, x2 = y1 + np.float32(0.3) + j[:, 2], x1 + np.float32(0.3) + j[:, 3]
    else:                                               # coordinates on a 1/8 grid: exact ties, IoU exactly at simple fractions, empty boxes
        g = rng.integers(0, 9, (n, 4)).astype(np.float32) / np.float32(8)
        y1, y2 = np.minimum(g[:, 0], g[:, 2]), np.maximum(g[:, 0], g[:, 2])
        x1, x2 = np.minimum(g[:, 1], g[:, 3]), np.maximum(g[:, 1], g[:, 3])
    return np.clip(np.stack([y1, x1, y2, x2], axis=1), 0, 1).astype(np.float32).reshape(n, 4)


@SET
@given(boxes(), st.sampled_from([0.3, 0.5, 0.7]), st.integers(1, 50))
def test_nms_literal_equals_oracle_and_is_idempotent(b, thr, max_keep):
    n = b.shape[0]
    keep = orc.nms(b, np.arange(n), thr, max_keep)
    ref = lit.non_max_supression(b.reshape(-1), range(n), thr, max_keep)
    assert keep.tolist() == ref
    again = orc.nms(b, keep, thr, max_keep)             # survivors do not suppress each other
    assert again.tolist() == keep.tolist()
    for i in keep:                                      # only boxes with positive area are ever selected (Utils.swift:195)
        assert b[i, 2] > b[i, 0] and b[i, 3] > b[i, 1]


@SET
@given(boxes(max_n=12))
def test_iou_literal_equals_oracle_and_is_symmetric(b):
    for i in range(b.shape[0]):
        for j in range(b.shape[0]):
            v = orc.iou(b[i], b[j])
            assert np.float32(v) == lit.IOU(lit.CGRect(b[i]), lit.CGRect(b[j]))
            assert v == orc.iou(b[j], b[i]) and 0.0 <= v <= 1.0


@SET
@given(boxes(), st.integers(0, 2 ** 31 - 1), st.floats(0.01, 3.0))
def test_box_decode_literal_equals_oracle(b, seed, scale):
    n = b.shape[0]
    d = (np.random.default_rng(seed).standard_normal((n, 4)) * scale).astype(np.float32)
    want = orc.clip(orc.apply_box_deltas(b, d)) if n else np.zeros((0, 4), np.float32)
    got = b.copy().reshape(-1)
    lit.apply_box_deltas(got, d.reshape(-1))
    got = lit.vDSP_vclip(got, 0.0, 1.0).reshape(n, 4)
    np.testing.assert_array_equal(got, want)
    assert ((got >= 0) & (got <= 1)).all()
    assert (got[:, 2] >= got[:, 0]).all() and (got[:, 3] >= got[:, 1]).all()     # exp > 0 and a monotone clamp keep the order


@SET
@given(boxes(), st.sampled_from([(1024.0, 1024.0), (512.0, 512.0), (640.0, 384.0)]))
def test_levels_literal_equal_oracle(b, size):
    w, h = size
    lv = orc.roi_levels(b, w, h) if b.shape[0] else np.zeros(0, np.int32)
    items = lit.rois_to_input_items(b, 224.0, w, h)
    np.testing.assert_array_equal(np.array([c[0] + 2 if c is not None else -1 for _, c in items], np.int32), lv)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(boxes(max_n=6), st.sampled_from([1, 2, 7, 14]), st.integers(0, 2 ** 31 - 1), st.sampled_from([(5, 9), (16, 16), (1, 1), (2, 3)]))
def test_crop_and_resize_literal_equals_oracle(b, pool, seed, hw):
    rng = np.random.default_rng(seed)
    fmap = rng.standard_normal((3, hw[0], hw[1])).astype(np.float32)
    maps = [fmap, fmap, fmap, fmap]
    if b.shape[0] == 0:
        return
    out, lv = orc.pyramid_roialign(b, maps, pool, 1024, 1024)
    for i in range(b.shape[0]):
        if lv[i] < 0:
            assert (out[i] == 0).all()
        else:
            np.testing.assert_array_equal(lit.crop_and_resize_bilinear(fmap, b[i], pool), out[i])
