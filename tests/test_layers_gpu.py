"""GPU parity tests: every custom layer through the C ABI (ctypes) against the CPU
oracle on the same seeded inputs.  Bit-exact for keep indices / class ids /
levels; boxes and ROIAlign values are compared exactly too (the arithmetic
contract makes them bit-identical; the north-star tolerance is 1e-4)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4   # north-star tolerance for boxes / masks / pooled features


@pytest.fixture(scope="module")
def anchors(pkg):
    return pkg.synth.generate_anchors(1024, 1024)


def test_anchor_count(anchors):
    assert anchors.shape == (261888, 4)


def _run_proposal(pkg, ctx, probs, deltas):
    b = probs.shape[0]
    mp = ctx.cfg.max_proposals
    rois = np.full((b, mp, 4), 7.0, np.float32)          # poisoned: every element must be written
    keep = np.full((b, mp), 99, np.int32)
    cnt = np.zeros(b, np.int32)
    pkg.ProposalLayer(context=ctx).evaluate([probs, deltas], [rois], keep_anchor=keep, count=cnt)
    return rois, keep, cnt


def test_proposal_full_size_batch(pkg, ctx, orc, anchors):
    ctx.set_anchors(anchors)
    imgs = [pkg.synth.rpn_outputs(anchors, i) for i in range(3)]
    probs = np.stack([p for p, _ in imgs])
    deltas = np.stack([d for _, d in imgs])
    rois, keep, cnt = _run_proposal(pkg, ctx, probs, deltas)
    for i in range(3):
        r0, k0, c0 = orc.proposal(probs[i], deltas[i], anchors)
        assert cnt[i] == c0
        np.testing.assert_array_equal(keep[i], k0)            # bit-exact NMS keep indices
        np.testing.assert_allclose(rois[i], r0, rtol=0, atol=TOL)
        np.testing.assert_array_equal(rois[i], r0)
        assert c0 == 1000                                     # the synthetic RPN fills all proposals


def test_proposal_ties_and_small_n(pkg, orc):
    # duplicated scores (tie order = lower anchor first), fewer anchors than pre_nms, <max survivors (Q4)
    c = pkg.Context(pre_nms_max_proposals=300, max_proposals=50)
    rng = np.random.default_rng(11)
    n = 1000
    a = pkg.synth.random_rois(n, 5, min_px=30, max_px=300)
    c.set_anchors(a)
    probs = np.zeros((1, n, 2), np.float32)
    probs[0, :, 1] = rng.integers(0, 20, n).astype(np.float32) / 20.0     # heavy ties
    probs[0, :, 0] = 1 - probs[0, :, 1]
    deltas = (rng.standard_normal((1, n, 4)) * 0.3).astype(np.float32)
    rois = np.zeros((1, 50, 4), np.float32); keep = np.zeros((1, 50), np.int32); cnt = np.zeros(1, np.int32)
    pkg.ProposalLayer({"preNMSMaxProposals": 300, "maxProposals": 50}, context=c).evaluate([probs, deltas], [rois], keep, cnt)
    r0, k0, c0 = orc.proposal(probs[0], deltas[0], a, pre_nms=300, max_proposals=50)
    assert cnt[0] == c0
    np.testing.assert_array_equal(keep[0], k0)
    np.testing.assert_array_equal(rois[0], r0)
    # all scores equal: the first 300 anchors by index are the candidates
    probs[0, :, 1] = 0.5
    pkg.ProposalLayer({"preNMSMaxProposals": 300, "maxProposals": 50}, context=c).evaluate([probs, deltas], [rois], keep, cnt)
    r0, k0, c0 = orc.proposal(probs[0], deltas[0], a, pre_nms=300, max_proposals=50)
    np.testing.assert_array_equal(keep[0], k0)
    assert k0[:c0].max() < 300
    # N < pre_nms and fewer survivors than max (the reference would trap here, Q4)
    c2 = pkg.Context(pre_nms_max_proposals=6000, max_proposals=1000)
    small = np.tile(np.array([[0.1, 0.1, 0.6, 0.6]], np.float32), (40, 1))
    small[:, 1] += np.arange(40, dtype=np.float32) * 1e-4
    c2.set_anchors(small)
    p = np.zeros((1, 40, 2), np.float32); p[0, :, 1] = np.linspace(0.1, 0.9, 40)
    d = np.zeros((1, 40, 4), np.float32)
    rois = np.ones((1, 1000, 4), np.float32); keep = np.zeros((1, 1000), np.int32); cnt = np.zeros(1, np.int32)
    pkg.ProposalLayer(context=c2).evaluate([p, d], [rois], keep, cnt)
    r0, k0, c0 = orc.proposal(p[0], d[0], small)
    assert cnt[0] == c0 == 1
    np.testing.assert_array_equal(keep[0], k0)
    np.testing.assert_array_equal(rois[0], r0)
    c.close(); c2.close()


def test_proposal_zero_area_boxes(pkg, orc):
    c = pkg.Context(pre_nms_max_proposals=64, max_proposals=16)
    a = pkg.synth.random_rois(64, 9, min_px=50, max_px=200)
    a[::3, 2] = a[::3, 0]                                   # zero height -> never selectable (Utils.swift:195)
    c.set_anchors(a)
    rng = np.random.default_rng(2)
    p = np.zeros((1, 64, 2), np.float32); p[0, :, 1] = rng.uniform(size=64)
    d = np.zeros((1, 64, 4), np.float32)
    rois = np.zeros((1, 16, 4), np.float32); keep = np.zeros((1, 16), np.int32); cnt = np.zeros(1, np.int32)
    pkg.ProposalLayer({"preNMSMaxProposals": 64, "maxProposals": 16}, context=c).evaluate([p, d], [rois], keep, cnt)
    r0, k0, c0 = orc.proposal(p[0], d[0], a, pre_nms=64, max_proposals=16)
    np.testing.assert_array_equal(keep[0], k0)
    np.testing.assert_array_equal(rois[0], r0)
    assert not np.isin(k0[:c0], np.arange(0, 64, 3)).any()
    c.close()


@pytest.mark.parametrize("pool,stride", [(7, 4), (14, 6)])
def test_pyramid_roialign_boundary_layout(pkg, ctx, orc, pool, stride):
    b, r, ch = 2, 200, 32
    maps = [np.stack([pkg.synth.feature_maps(i, channels=ch)[l] for i in range(b)]) for l in range(4)]
    rois = np.zeros((b, r, stride), np.float32)
    for i in range(b):
        rois[i, :, :4] = pkg.synth.random_rois(r, i, n_pad=7)
        rois[i, 5, :4] = [0.2, 0.3, 0.2, 0.5]                 # zero-height roi -> padding block
    out = np.full((b, r, ch, pool, pool), 3.0, np.float32)
    lv = np.zeros((b, r), np.int32)
    pkg.PyramidROIAlignLayer({"poolSize": pool}, context=ctx).evaluate([rois] + maps, [out], level=lv)
    for i in range(b):
        o0, l0 = orc.pyramid_roialign(rois[i], [m[i] for m in maps], pool)
        np.testing.assert_array_equal(lv[i], l0)
        np.testing.assert_allclose(out[i], o0, rtol=0, atol=TOL)
        np.testing.assert_array_equal(out[i], o0)
        assert set(np.unique(l0)) >= {-1, 2, 3, 4, 5}
        assert not out[i, r - 7:].any() and not out[i, 5].any()


def test_pyramid_roialign_nhwc_f16(pkg, ctx, orc):
    import ctypes as C
    import torch
    b, r, ch, pool = 2, 150, 256, 7
    maps = [np.stack([pkg.synth.feature_maps(10 + i, channels=ch)[l] for i in range(b)]) for l in range(4)]
    hwc = [np.ascontiguousarray(m.transpose(0, 2, 3, 1)).astype(np.float16) for m in maps]
    rois = np.stack([pkg.synth.random_rois(r, 20 + i, n_pad=3) for i in range(b)])
    d_maps = [torch.from_numpy(m).cuda() for m in hwc]
    d_rois = torch.from_numpy(rois).cuda()
    d_out = torch.full((b, r, pool, pool, ch), 5.0, dtype=torch.float16, device="cuda")
    d_lv = torch.zeros((b, r), dtype=torch.int32, device="cuda")
    hw = (C.c_int32 * 8)(*[d for m in hwc for d in m.shape[1:3]])
    fp = (C.c_void_p * 4)(*[m.data_ptr() for m in d_maps])
    pkg._cabi.check(ctx.handle, pkg.lib().mrcnn_roialign_nhwc_f16(ctx.handle, b, d_rois.data_ptr(), 4, r, fp, hw, ch, pool,
                                                               d_out.data_ptr(), d_lv.data_ptr()))
    ctx.synchronize()
    out = d_out.cpu().numpy(); lv = d_lv.cpu().numpy()
    for i in range(b):
        o0, l0 = orc.pyramid_roialign_nhwc_f16(rois[i], [m[i] for m in hwc], pool)
        np.testing.assert_array_equal(lv[i], l0)
        np.testing.assert_array_equal(out[i].view(np.uint16), o0.view(np.uint16))      # bit-exact fp16


@pytest.mark.parametrize("pool,ch", [(7, 256), (14, 64), (7, 32), (3, 40), (16, 32), (20, 16), (7, 12)])
def test_roialign_boundary_layout_every_path(pkg, ctx, orc, pool, ch):
    """Boundary (CHW fp32) layout: the staged kernel's ring / gather / padding paths, run-time pool sizes, channel counts
    TMA boxes cannot cover (12: the plain gather kernel), bit for bit against the oracle."""
    b, r = 2, 120
    maps = [np.stack([pkg.synth.feature_maps(90 + i, channels=ch)[l] for i in range(b)]) for l in range(4)]
    rois = np.stack([_adversarial_rois(pkg, r, 95 + i) for i in range(b)])
    out = np.full((b, r, ch, pool, pool), 3.0, np.float32)
    lv = np.zeros((b, r), np.int32)
    pkg.PyramidROIAlignLayer({"poolSize": pool}, context=ctx).evaluate([rois] + maps, [out], level=lv)
    for i in range(b):
        o0, l0 = orc.pyramid_roialign(rois[i], [m[i] for m in maps], pool)
        np.testing.assert_array_equal(lv[i], l0)
        np.testing.assert_array_equal(out[i].view(np.uint32), o0.view(np.uint32))
    assert {-1, 2, 3, 4, 5} <= set(np.unique(lv))


def test_roialign_boundary_layout_gather_variant_agrees(pkg, orc, monkeypatch):
    monkeypatch.setenv("MRCNN_ROIALIGN", "gather")
    c = pkg.Context()
    try:
        b, r, ch, pool = 1, 120, 32, 7
        maps = [np.stack([pkg.synth.feature_maps(97, channels=ch)[l]]) for l in range(4)]
        rois = np.stack([_adversarial_rois(pkg, r, 98)])
        out = np.zeros((b, r, ch, pool, pool), np.float32)
        pkg.PyramidROIAlignLayer({"poolSize": pool}, context=c).evaluate([rois] + maps, [out])
        o0, _ = orc.pyramid_roialign(rois[0], [m[0] for m in maps], pool)
        np.testing.assert_array_equal(out[0].view(np.uint32), o0.view(np.uint32))
    finally:
        c.close()


def _run_nhwc(pkg, ctx, rois, hwc, pool, stride=4):
    import ctypes as C
    import torch
    b, r = rois.shape[:2]
    ch = hwc[0].shape[-1]
    d_maps = [torch.from_numpy(m).cuda() for m in hwc]
    d_rois = torch.from_numpy(rois).cuda()
    d_out = torch.full((b, r, pool, pool, ch), 5.0, dtype=torch.float16, device="cuda")
    d_lv = torch.zeros((b, r), dtype=torch.int32, device="cuda")
    hw = (C.c_int32 * 8)(*[d for m in hwc for d in m.shape[1:3]])
    fp = (C.c_void_p * 4)(*[m.data_ptr() for m in d_maps])
    pkg._cabi.check(ctx.handle, pkg.lib().mrcnn_roialign_nhwc_f16(ctx.handle, b, d_rois.data_ptr(), stride, r, fp, hw, ch, pool,
                                                               d_out.data_ptr(), d_lv.data_ptr()))
    ctx.synchronize()
    return d_out.cpu().numpy(), d_lv.cpu().numpy()


def _adversarial_rois(pkg, r, seed):
    """Random rois plus the cases every branch of the staged kernel (roialign.cu) has to get right: boxes touching the
    image border (last sample lands on / one ulp beyond the last pixel -> TF's extrapolation value 0), boxes reaching
    outside [0, 1] (leading / trailing out-of-range samples), inverted boxes (the ring cannot serve them: gather path),
    boxes wider than a ring slot (level 5, > 32 feature pixels is impossible at 1024^2, so a wide level-2 box via tiny
    height), degenerate boxes (padding), a sample row exactly on a feature row, tiny boxes (all samples in one pixel)."""
    rois = pkg.synth.random_rois(r, seed, n_pad=3).copy()
    rng = np.random.default_rng(seed)
    special = [
        [0.0, 0.0, 1.0, 1.0], [0.5, 0.5, 1.0, 1.0], [0.0, 0.0, 0.03, 0.05], [0.9, 0.9, 1.0, 1.0],
        [0.25, 0.25, 0.25 + 6 / 255.0, 0.25 + 6 / 255.0],            # samples on feature rows of P2
        [0.3, 0.3, 0.3001, 0.3001],                                     # all samples inside one pixel
        [-0.1, -0.05, 0.4, 0.5], [0.6, 0.7, 1.2, 1.1], [-0.2, -0.2, 1.3, 1.3],      # outside the map on one / both ends
        [0.8, 0.6, 0.2, 0.1],                                           # inverted on both axes: valid level, gather path
        [0.4, 0.1, 0.401, 0.95],                                        # very wide, very flat: level 2, > 32 pixels wide
        [0.2, 0.2, 0.2, 0.6], [0.2, 0.6, 0.5, 0.2],                     # zero / negative area -> padding
        [np.nan, 0.1, 0.5, 0.5],
    ]
    for k, sp in enumerate(special):
        rois[2 * k + 1, :4] = sp
    for k in range(20):                                                  # boxes clipped at 1.0 like ProposalLayer's output
        y1, x1 = rng.uniform(0.3, 0.95, 2)
        rois[40 + k, :4] = [y1, x1, 1.0 if k % 2 else min(1.0, y1 + 0.2), 1.0 if k % 3 else min(1.0, x1 + 0.3)]
    return rois.astype(np.float32)


@pytest.mark.parametrize("pool,ch", [(7, 256), (14, 256), (7, 32), (14, 64), (1, 256), (2, 256), (5, 128), (16, 256), (20, 64)])
def test_roialign_nhwc_f16_every_path(pkg, ctx, orc, pool, ch):
    """The staged kernel against the oracle, bit for bit, on rois that drive each of its paths: the predicate-free ring
    loop (pool 7, 256 channels), the generic ring loop (out-of-range samples, fewer channels), the row-wise loop (pool
    14), run-time pool sizes (1, 2, 5, 16), the gather path (inverted / too wide boxes, pool > 16) and padding blocks."""
    b, r = 2, 120
    maps = [np.stack([pkg.synth.feature_maps(30 + i, channels=ch)[l] for i in range(b)]) for l in range(4)]
    hwc = [np.ascontiguousarray(m.transpose(0, 2, 3, 1)).astype(np.float16) for m in maps]
    rois = np.stack([_adversarial_rois(pkg, r, 40 + i) for i in range(b)])
    out, lv = _run_nhwc(pkg, ctx, rois, hwc, pool)
    for i in range(b):
        o0, l0 = orc.pyramid_roialign_nhwc_f16(rois[i], [m[i] for m in hwc], pool)
        np.testing.assert_array_equal(lv[i], l0)
        np.testing.assert_array_equal(out[i].view(np.uint16), o0.view(np.uint16))
    assert {-1, 2, 3, 4, 5} <= set(np.unique(lv))


@pytest.mark.parametrize("env", [{"MRCNN_ROIALIGN": "gather"}, {"MRCNN_ROIALIGN_ROWWISE": "1"}, {"MRCNN_ROIALIGN_ROWWISE": "0"},
                                 {"MRCNN_ROIALIGN_CTAS": "3"}, {"MRCNN_ROIALIGN_CTAS": "2"}, {"MRCNN_ROIALIGN_SLOT_PX": "8"}, {"MRCNN_ROIALIGN_SORT": "1"}])
def test_roialign_nhwc_f16_kernel_variants_agree(pkg, orc, env, monkeypatch):
    """Every variant of the kernel (pure gather, either consumer loop for either pool size, three CTAs per SM, narrow ring
    slots = most rois on the gather path, sorted processing order) gives the oracle's bits.  The knobs are read once per context."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    c = pkg.Context()
    try:
        b, r, ch = 2, 120, 256
        maps = [np.stack([pkg.synth.feature_maps(50 + i, channels=ch)[l] for i in range(b)]) for l in range(4)]
        hwc = [np.ascontiguousarray(m.transpose(0, 2, 3, 1)).astype(np.float16) for m in maps]
        rois = np.stack([_adversarial_rois(pkg, r, 60 + i) for i in range(b)])
        for pool in (7, 14):
            out, lv = _run_nhwc(pkg, c, rois, hwc, pool)
            for i in range(b):
                o0, l0 = orc.pyramid_roialign_nhwc_f16(rois[i], [m[i] for m in hwc], pool)
                np.testing.assert_array_equal(lv[i], l0)
                np.testing.assert_array_equal(out[i].view(np.uint16), o0.view(np.uint16))
    finally:
        c.close()


def test_roialign_nhwc_f16_many_rois_per_cta(pkg, ctx, orc):
    """More rois than the persistent grid has CTAs (2 x 148): descriptors and ring entries wrap many times, rois are
    handed out by the ticket counter; repeated launches re-arm it."""
    b, r, ch, pool = 3, 1000, 256, 7
    maps = [np.stack([pkg.synth.feature_maps(70 + i, channels=ch)[l] for i in range(b)]) for l in range(4)]
    hwc = [np.ascontiguousarray(m.transpose(0, 2, 3, 1)).astype(np.float16) for m in maps]
    rois = np.stack([pkg.synth.random_rois(r, 80 + i, n_pad=11) for i in range(b)])
    for _ in range(2):
        out, lv = _run_nhwc(pkg, ctx, rois, hwc, pool)
    for i in range(b):
        o0, l0 = orc.pyramid_roialign_nhwc_f16(rois[i], [m[i] for m in hwc], pool)
        np.testing.assert_array_equal(lv[i], l0)
        np.testing.assert_array_equal(out[i].view(np.uint16), o0.view(np.uint16))


def test_classifier_select(pkg, ctx, orc):
    probs = np.stack([pkg.synth.classifier_outputs(500, i)[0] for i in range(2)])
    bbox = np.stack([pkg.synth.classifier_outputs(500, i)[1] for i in range(2)])
    probs[0, 3, 5] = probs[0, 3, 9] = 0.99                     # tie -> first max (Q7)
    out = np.zeros((2, 500, 6), np.float32)
    pkg.TimeDistributedClassifierLayer(context=ctx).select(probs, bbox, out)
    for i in range(2):
        np.testing.assert_array_equal(out[i], orc.classifier_select(probs[i], bbox[i]))
    assert out[0, 3, 4] == 5


def _cls_from_synth(pkg, orc, r, seed):
    p, bb = pkg.synth.classifier_outputs(r, seed)
    return orc.classifier_select(p, bb)


def test_detection_layer(pkg, ctx, orc):
    b, r = 3, 1000
    rois = np.stack([pkg.synth.random_rois(r, 40 + i, min_px=20, max_px=500) for i in range(b)])
    cls = np.stack([_cls_from_synth(pkg, orc, r, 40 + i) for i in range(b)])
    cls[..., :4] *= 0.5
    # image 2: heavy clustering so per-class NMS really suppresses, plus score ties
    rois[2, 100:400] = rois[2, 100] + (np.random.default_rng(3).uniform(-0.01, 0.01, (300, 4))).astype(np.float32)
    rois[2] = np.clip(rois[2], 0, 1)
    cls[2, 100:400, 4] = 3.0
    cls[2, 100:400, 5] = np.float32(0.9)
    out = np.full((b, 100, 6), 9.0, np.float32); keep = np.zeros((b, 100), np.int32); cnt = np.zeros(b, np.int32)
    pkg.DetectionLayer(context=ctx).evaluate([rois, cls], [out], keep_roi=keep, count=cnt)
    for i in range(b):
        o0, k0, c0 = orc.detection(rois[i], cls[i])
        assert cnt[i] == c0
        np.testing.assert_array_equal(keep[i], k0)             # bit-exact keep indices
        np.testing.assert_array_equal(out[i, :, 4], o0[:, 4])  # bit-exact class ids
        np.testing.assert_allclose(out[i], o0, rtol=0, atol=TOL)
        np.testing.assert_array_equal(out[i], o0)
    assert cnt.max() == 100


def test_detection_layer_edges(pkg, ctx, orc):
    r = 64
    rois = pkg.synth.random_rois(r, 77)[None]
    cls = np.zeros((1, r, 6), np.float32)
    out = np.ones((1, 100, 6), np.float32); keep = np.zeros((1, 100), np.int32); cnt = np.ones(1, np.int32)
    pkg.DetectionLayer(context=ctx).evaluate([rois, cls], [out], keep, cnt)     # nothing passes
    assert cnt[0] == 0 and not out.any() and (keep == -1).all()
    cls[0, :, 4] = 1; cls[0, :, 5] = np.float32(0.7)                             # == threshold: kept (Q8)
    cls[0, 10, 5] = np.nextafter(np.float32(0.7), np.float32(0))
    pkg.DetectionLayer(context=ctx).evaluate([rois, cls], [out], keep, cnt)
    o0, k0, c0 = orc.detection(rois[0], cls[0])
    assert cnt[0] == c0 and 10 not in keep[0]
    np.testing.assert_array_equal(keep[0], k0)
    np.testing.assert_array_equal(out[0], o0)


def test_detections_decode(pkg, ctx, orc):
    import ctypes as C
    rng = np.random.default_rng(8)
    det = np.zeros((2, 100, 6), np.float32)
    det[:, :30, :4] = np.sort(rng.uniform(size=(2, 30, 4)).astype(np.float32), axis=-1)[..., [0, 1, 2, 3]]
    det[:, :30, 4] = rng.integers(1, 81, (2, 30))
    det[:, :30, 5] = rng.uniform(0.6, 1.0, (2, 30))
    det[0, 3, 5] = np.float32(0.7)
    masks = rng.uniform(size=(2, 100, 28, 28)).astype(np.float32)
    cnt = np.zeros(2, np.int32); idx = np.zeros((2, 100), np.int32); bbox = np.zeros((2, 100, 4), np.float64)
    cls = np.zeros((2, 100), np.int32); score = np.zeros((2, 100), np.float64); mu8 = np.zeros((2, 100, 784), np.uint8)
    P = pkg._cabi.ptr
    pkg._cabi.check(ctx.handle, pkg.lib().mrcnn_detections_decode(ctx.handle, 2, P(det), P(masks), P(cnt), P(idx), P(bbox),
                                                               P(cls), P(score), P(mu8)))
    for i in range(2):
        n0, i0, b0, c0, s0, m0 = orc.detections_decode(det[i], masks[i])
        assert cnt[i] == n0
        np.testing.assert_array_equal(idx[i, :n0], i0[:n0]); np.testing.assert_array_equal(cls[i, :n0], c0[:n0])
        np.testing.assert_array_equal(bbox[i, :n0], b0[:n0]); np.testing.assert_array_equal(score[i, :n0], s0[:n0])
        np.testing.assert_array_equal(mu8[i, :n0], m0[:n0])


def test_error_paths(pkg, ctx):
    rois = np.zeros((1, 1000, 4), np.float32)
    with pytest.raises(pkg.MaskRCNNError):
        # anchors for a different N
        ctx.set_anchors(np.zeros((10, 4), np.float32))
        pkg.ProposalLayer(context=ctx).evaluate([np.zeros((1, 20, 2), np.float32), np.zeros((1, 20, 4), np.float32)], [rois])
    with pytest.raises(pkg.MaskRCNNError):
        pkg.ProposalLayer({"maxProposals": 5}, context=ctx)     # parameters must match the context
    with pytest.raises(pkg.MaskRCNNError):
        pkg.Context(anchors_path="/nonexistent/anchors.bin")


def test_letterbox_matches_oracle(pkg, orc):
    c = pkg.Context(image_h=256, image_w=256)
    rng = np.random.default_rng(5)
    for h, w in ((100, 180), (333, 211), (256, 256), (37, 500)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out = np.full((256, 256, 3), 7, np.uint8)
        pkg._cabi.check(c.handle, pkg.lib().mrcnn_letterbox_eval(c.handle, pkg._cabi.ptr(img), h, w, pkg._cabi.ptr(out)))
        np.testing.assert_array_equal(out, orc.letterbox(img, 256, 256))           # bit-exact (same fp64 op order)
    c.close()


@pytest.mark.parametrize("seed,n,pre,mx,thr", [(1, 3000, 700, 120, 0.7), (2, 10000, 1500, 300, 0.5), (3, 777, 777, 64, 0.9),
                                               (4, 50000, 6000, 1000, 0.7), (5, 4097, 2048, 2048, 0.3), (6, 129, 65, 65, 0.6)])
def test_proposal_randomised_configs(pkg, orc, seed, n, pre, mx, thr):
    """Odd sizes: N not a multiple of the radix tile, pre_nms at / across 64- and 256-box boundaries, max > survivors."""
    rng = np.random.default_rng(seed)
    c = pkg.Context(pre_nms_max_proposals=pre, max_proposals=mx, proposal_nms_iou=float(np.float32(thr)))
    a = pkg.synth.random_rois(n, seed, min_px=10, max_px=400)
    c.set_anchors(a)
    p = np.zeros((1, n, 2), np.float32)
    p[0, :, 1] = np.round(rng.uniform(size=n) * 512) / 512          # many exact ties
    p[0, :, 0] = 1 - p[0, :, 1]
    d = (rng.standard_normal((1, n, 4)) * 0.5).astype(np.float32)
    rois = np.full((1, mx, 4), 3.0, np.float32); keep = np.zeros((1, mx), np.int32); cnt = np.zeros(1, np.int32)
    pkg.ProposalLayer({"preNMSMaxProposals": pre, "maxProposals": mx, "nmsIOUThreshold": float(np.float32(thr))}, context=c).evaluate(
        [p, d], [rois], keep, cnt)
    r0, k0, c0 = orc.proposal(p[0], d[0], a, pre_nms=pre, max_proposals=mx, iou_thr=thr)
    assert cnt[0] == c0
    np.testing.assert_array_equal(keep[0], k0)
    np.testing.assert_array_equal(rois[0], r0)
    c.close()


@pytest.mark.parametrize("seed,r,ncls,maxdet", [(1, 300, 81, 100), (2, 1024, 81, 100), (3, 2000, 81, 100), (4, 1500, 5, 40),
                                                (5, 65, 300, 10), (6, 4000, 81, 100)])
def test_detection_randomised_configs(pkg, orc, seed, r, ncls, maxdet):
    """Exercises the register path (<= 1024 rois), the shared-memory path and the generic sequential fallback
    (bitmaps larger than shared memory) of the per-class NMS, few / many classes and small per-class caps."""
    rng = np.random.default_rng(100 + seed)
    c = pkg.Context(num_classes=ncls, max_detections=maxdet)
    rois = pkg.synth.random_rois(r, seed, min_px=20, max_px=300)
    centres = rng.integers(0, 12, r)
    rois = np.clip(rois * 0.2 + (centres[:, None] / 16.0 + 0.05), 0, 1).astype(np.float32)       # 12 heavy clusters
    cls = np.zeros((1, r, 6), np.float32)
    cls[0, :, :4] = (rng.standard_normal((r, 4)) * 0.3).astype(np.float32)
    cls[0, :, 4] = rng.integers(0, min(ncls, 9), r)                  # class 0 = background is filtered
    cls[0, :, 5] = np.round(rng.uniform(0.5, 1.0, r) * 64) / 64      # ties + below-threshold rows
    out = np.full((1, maxdet, 6), 9.0, np.float32); keep = np.zeros((1, maxdet), np.int32); cnt = np.zeros(1, np.int32)
    pkg.DetectionLayer({"maxDetections": maxdet}, context=c).evaluate([rois[None], cls], [out], keep, cnt)
    o0, k0, c0 = orc.detection(rois, cls[0], max_det=maxdet)
    assert cnt[0] == c0
    np.testing.assert_array_equal(keep[0], k0)
    np.testing.assert_array_equal(out[0], o0)
    c.close()


def test_detection_layer_inverted_rois_follow_cgrect_standardisation(pkg, ctx, orc):
    """DetectionLayer on caller-supplied rois with x2 < x1 / y2 < y1: the decoded boxes stay inverted, and the NMS treats
    them as CGRect does (|w|, |h|, min / max edges, Utils.swift:194-246) -- the fp32 fast path of the bitmask kernel
    must not be taken for them.  Bit-exact against the oracle."""
    rng = np.random.default_rng(12)
    r = 600
    rois = pkg.synth.random_rois(r, 33, min_px=40, max_px=300).copy()
    inv = rng.random(r) < 0.35
    rois[inv] = rois[inv][:, [2, 3, 0, 1]]                          # both axes inverted: valid level, negative w and h
    pr, bb = pkg.synth.classifier_outputs(r, 34)
    cls = orc.classifier_select(pr, bb)
    cls[:, :4] *= np.float32(0.1)                                   # small refinements: the inversion survives the decode
    cls[:, 4] = np.where(cls[:, 4] > 0, 1 + (cls[:, 4] % 3), 0)     # few classes -> plenty of same-class overlaps
    det = np.zeros((1, 100, 6), np.float32); keep = np.zeros((1, 100), np.int32); cnt = np.zeros(1, np.int32)
    pkg.DetectionLayer(context=ctx).evaluate([rois[None], cls[None]], [det], keep, cnt)
    d0, k0, n0 = orc.detection(rois, cls)
    assert cnt[0] == n0 and n0 > 5
    np.testing.assert_array_equal(keep[0, :n0], k0[:n0])
    np.testing.assert_array_equal(det[0], d0)
    b = det[0, :n0, :4]
    assert ((b[:, 2] < b[:, 0]) | (b[:, 3] < b[:, 1])).any()        # inverted boxes are among the detections


def test_roialign_abi_rejects_bad_arguments_without_poisoning_the_context(pkg, ctx, orc):
    """A NULL level pointer, a zero map size or a pool size out of range are refused with MRCNN_EINVAL before anything is
    staged or launched (a NULL map dereferenced on the device would be a sticky fault); the context keeps working."""
    import ctypes as C
    import torch
    maps = [torch.zeros((1, s, s, 32), dtype=torch.float16, device="cuda") for s in (64, 32, 16, 8)]
    rois = torch.from_numpy(pkg.synth.random_rois(16, 1)[None]).cuda()
    out = torch.zeros((1, 16, 7, 7, 32), dtype=torch.float16, device="cuda")
    hw = (C.c_int32 * 8)(64, 64, 32, 32, 16, 16, 8, 8)
    lib = pkg.lib()
    def call(fp, hwv=hw, pool=7, ch=32):
        return lib.mrcnn_roialign_nhwc_f16(ctx.handle, 1, rois.data_ptr(), 4, 16, fp, hwv, ch, pool, out.data_ptr(), None)
    good = (C.c_void_p * 4)(*[m.data_ptr() for m in maps])
    bad = (C.c_void_p * 4)(maps[0].data_ptr(), None, maps[2].data_ptr(), maps[3].data_ptr())
    assert call(bad) == pkg._cabi.EINVAL and b"null feature map" in lib.mrcnn_last_error(ctx.handle)
    assert call(good, (C.c_int32 * 8)(64, 64, 0, 32, 16, 16, 8, 8)) == pkg._cabi.EINVAL
    assert call(good, pool=65) == pkg._cabi.EINVAL and call(good, pool=0) == pkg._cabi.EINVAL
    assert call(good, ch=36) == pkg._cabi.EINVAL                     # channels must be a multiple of 8
    fp32 = [torch.zeros((1, 8, s, s), device="cuda") for s in (64, 32, 16, 8)]
    o32 = torch.zeros((1, 16, 8, 7, 7), device="cuda")
    badf = (C.c_void_p * 4)(fp32[0].data_ptr(), fp32[1].data_ptr(), None, fp32[3].data_ptr())
    assert lib.mrcnn_pyramid_roialign_eval(ctx.handle, 1, rois.data_ptr(), 4, 16, badf, hw, 8, 7, o32.data_ptr(), None) == pkg._cabi.EINVAL
    assert call(good) == 0                                           # the context is still usable
    ctx.synchronize()
    assert not out.cpu().numpy().any()                               # zero maps -> zero pooled features
