"""CPU tests: the oracle against hand-worked known answers (the reference ships no
golden vectors -- SURVEY.md section 4 -- so these are authored from the Swift
semantics, Appendix A) and against the committed golden fixtures."""
import numpy as np
import pytest


def test_iou_known_answers(orc):
    a = [0.0, 0.0, 0.5, 0.5]
    assert orc.iou(a, a) == pytest.approx(1.0)
    assert orc.iou(a, [0.5, 0.5, 1.0, 1.0]) == 0.0
    # half overlap: inter 0.125, union 0.375
    assert orc.iou(a, [0.0, 0.25, 0.5, 0.75]) == np.float32(0.125 / 0.375)
    # zero-area box -> 0 (Utils.swift:234-238)
    assert orc.iou(a, [0.2, 0.2, 0.2, 0.4]) == 0.0
    assert orc.iou([0.2, 0.2, 0.2, 0.4], a) == 0.0


def test_iou_is_double_then_float(orc):
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = np.sort(rng.uniform(0, 1, 4).astype(np.float32).reshape(2, 2), axis=0).T.reshape(-1)[[0, 2, 1, 3]]
        b = np.sort(rng.uniform(0, 1, 4).astype(np.float32).reshape(2, 2), axis=0).T.reshape(-1)[[0, 2, 1, 3]]
        ad, bd = a.astype(np.float64), b.astype(np.float64)
        ih = max(min(ad[2], bd[2]) - max(ad[0], bd[0]), 0.0)
        iw = max(min(ad[3], bd[3]) - max(ad[1], bd[1]), 0.0)
        inter = ih * iw
        ua = (ad[2] - ad[0]) * (ad[3] - ad[1]) + (bd[2] - bd[0]) * (bd[3] - bd[1]) - inter
        want = np.float32(inter / ua) if ua > 0 and (ad[2] - ad[0]) * (ad[3] - ad[1]) > 0 and (bd[2] - bd[0]) * (bd[3] - bd[1]) > 0 else np.float32(0)
        assert orc.iou(a, b) == want


def test_box_deltas_known_answer(orc):
    # zero deltas leave the box unchanged (up to fp32 rounding of the centre form)
    box = np.array([[0.25, 0.25, 0.75, 0.75]], dtype=np.float32)
    out = orc.apply_box_deltas(box, np.zeros((1, 4), np.float32))
    np.testing.assert_array_equal(out, box)
    # dy = 0.5 moves the centre by half the height; log dh = ln 2 doubles the height
    d = np.array([[0.5, 0.0, np.log(2.0), 0.0]], dtype=np.float32)
    out = orc.apply_box_deltas(box, d)
    np.testing.assert_allclose(out, [[0.25, 0.25, 1.25, 0.75]], rtol=0, atol=2e-7)
    np.testing.assert_allclose(orc.clip(out), [[0.25, 0.25, 1.0, 0.75]], rtol=0, atol=2e-7)


def test_nms_three_boxes(orc):
    boxes = np.array([[0.0, 0.0, 0.5, 0.5], [0.0, 0.02, 0.5, 0.52], [0.5, 0.5, 1.0, 1.0]], dtype=np.float32)
    np.testing.assert_array_equal(orc.nms(boxes, [0, 1, 2], 0.7, 10), [0, 2])
    np.testing.assert_array_equal(orc.nms(boxes, [1, 0, 2], 0.7, 10), [1, 2])      # order dependent
    np.testing.assert_array_equal(orc.nms(boxes, [0, 1, 2], 0.95, 10), [0, 1, 2])
    np.testing.assert_array_equal(orc.nms(boxes, [0, 1, 2], 0.7, 1), [0])           # max (Utils.swift:192)
    z = np.array([[0.1, 0.1, 0.1, 0.3], [0.0, 0.0, 0.5, 0.5]], dtype=np.float32)
    np.testing.assert_array_equal(orc.nms(z, [0, 1], 0.7, 10), [1])                 # zero area skipped (:195)


def test_iou_exactly_at_threshold_is_kept(orc):
    # IoU == thr is NOT suppressed: strict '>' (Utils.swift:203)
    a = np.array([0.0, 0.0, 1.0, 1.0], np.float32)
    b = np.array([0.0, 0.0, 1.0, 0.5], np.float32)     # IoU = 0.5 exactly
    boxes = np.stack([a, b])
    np.testing.assert_array_equal(orc.nms(boxes, [0, 1], 0.5, 10), [0, 1])
    np.testing.assert_array_equal(orc.nms(boxes, [0, 1], np.nextafter(np.float32(0.5), np.float32(0)), 10), [0])


def test_argsort_ties_lower_index_first(orc):
    k = np.array([0.5, 0.9, 0.5, 0.9, 0.1], np.float32)
    np.testing.assert_array_equal(orc.argsort_desc(k), [1, 3, 0, 2, 4])
    rng = np.random.default_rng(0)
    k = rng.integers(0, 50, 5000).astype(np.float32)
    want = np.lexsort((np.arange(k.size), -k))
    np.testing.assert_array_equal(orc.argsort_desc(k), want)


def test_roi_levels(orc):
    # sqrt(w*h) = 224/1024 -> level 4; x2 -> 5; /2 -> 3; tiny -> clamp 2; huge -> clamp 5; zero area -> padding
    s = 224.0 / 1024.0
    rois = np.array([[0, 0, s, s], [0, 0, 2 * s, 2 * s], [0, 0, s / 2, s / 2], [0, 0, 0.001, 0.001],
                     [0, 0, 1, 1], [0.3, 0.3, 0.3, 0.6], [0, 0, 0, 0]], dtype=np.float32)
    np.testing.assert_array_equal(orc.roi_levels(rois), [4, 5, 3, 2, 5, -1, -1])
    # image size enters through sqrt(W*H): 512x512 shifts everything by one level (Q15 intended)
    np.testing.assert_array_equal(orc.roi_levels(rois[:3], 512, 512), [3, 4, 2])


def test_crop_and_resize_known_answer(orc):
    # a 2x2 ramp map sampled over the whole map with pool 3: corners exact, centre = mean
    c = 4
    maps = [np.zeros((c, 2, 2), np.float32) for _ in range(4)]
    for m in maps:
        m[:, 0, 0], m[:, 0, 1], m[:, 1, 0], m[:, 1, 1] = 1.0, 2.0, 3.0, 4.0
    rois = np.array([[0.0, 0.0, 1.0, 1.0]], np.float32)      # level clamps to 5 -> maps[3]
    out, lv = orc.pyramid_roialign(rois, maps, 3)
    assert lv[0] == 5
    want = np.array([[1.0, 1.5, 2.0], [2.0, 2.5, 3.0], [3.0, 3.5, 4.0]], np.float32)
    for ch in range(c):
        np.testing.assert_array_equal(out[0, ch], want)
    # padding roi -> zero block (copyOutput :265-271, Q5 fixed)
    out, lv = orc.pyramid_roialign(np.zeros((2, 4), np.float32), maps, 3)
    assert (lv == -1).all() and not out.any()


def test_crop_and_resize_matches_grid_sample(orc):
    torch = pytest.importorskip("torch")
    import torch.nn.functional as F
    rng = np.random.default_rng(3)
    maps = [rng.standard_normal((8, s, s)).astype(np.float32) for s in (64, 32, 16, 8)]
    rois = np.array([[0.1, 0.2, 0.33, 0.41], [0.5, 0.5, 0.9, 0.95], [0.0, 0.0, 1.0, 1.0], [0.31, 0.02, 0.36, 0.09]], np.float32)
    out, lv = orc.pyramid_roialign(rois, maps, 7, image_w=256, image_h=256)
    for i, r in enumerate(rois):
        m = torch.from_numpy(maps[lv[i] - 2])[None]
        ys = torch.linspace(float(r[0]), float(r[2]), 7, dtype=torch.float64) * 2 - 1
        xs = torch.linspace(float(r[1]), float(r[3]), 7, dtype=torch.float64) * 2 - 1
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        grid = torch.stack([gx, gy], dim=-1)[None].float()
        ref = F.grid_sample(m, grid, mode="bilinear", align_corners=True)[0].numpy()
        np.testing.assert_allclose(out[i], ref, rtol=0, atol=2e-5)


def test_classifier_select_first_max(orc):
    p = np.array([[0.2, 0.5, 0.5, 0.1], [0.9, 0.05, 0.05, 0.0]], np.float32)
    b = np.arange(2 * 16, dtype=np.float32).reshape(2, 16)
    out = orc.classifier_select(p, b)
    np.testing.assert_array_equal(out[0], [4, 5, 6, 7, 1, 0.5])      # first max wins (Q7)
    np.testing.assert_array_equal(out[1], [16, 17, 18, 19, 0, np.float32(0.9)])


def test_detection_layer_semantics(orc):
    # 4 rois: two overlapping of class 3, one of class 5, one background; deltas zero
    rois = np.array([[0.1, 0.1, 0.5, 0.5], [0.1, 0.12, 0.5, 0.52], [0.1, 0.1, 0.5, 0.5], [0.6, 0.6, 0.9, 0.9]], np.float32)
    cls = np.zeros((4, 6), np.float32)
    cls[:, 4] = [3, 3, 5, 0]
    cls[:, 5] = [0.8, 0.95, 0.75, 0.99]
    out, keep, cnt = orc.detection(rois, cls)
    # NMS visits in ROI order (Q10): roi 0 suppresses roi 1 although roi 1 scores higher;
    # background dropped; result sorted by score
    assert cnt == 2
    np.testing.assert_array_equal(keep[:2], [0, 2])
    np.testing.assert_array_equal(out[0, 4:], [3, np.float32(0.8)])
    np.testing.assert_array_equal(out[1, 4:], [5, np.float32(0.75)])
    assert not out[2:].any() and (keep[2:] == -1).all()
    # score exactly at the Float threshold is kept (>=, Q8); just below is dropped
    cls[:, 5] = [np.float32(0.7), np.nextafter(np.float32(0.7), np.float32(0)), 0.75, 0.99]
    out, keep, cnt = orc.detection(rois, cls)
    np.testing.assert_array_equal(keep[:cnt], [2, 0])


def test_detections_decode(orc):
    det = np.zeros((3, 6), np.float32)
    det[0] = [0.1, 0.2, 0.5, 0.8, 7, 0.9]
    det[1] = [0.1, 0.2, 0.5, 0.8, 2, np.float32(0.7)]      # Float(0.7) fails the Double > 0.7 test (Q8)
    det[2] = [0.0, 0.0, 1.0, 1.0, 1, 0.71]
    masks = np.zeros((3, 28, 28), np.float32)
    masks[0] = 1.0
    masks[2] = 0.5
    n, idx, bbox, cls, score, mu8 = orc.detections_decode(det, masks)
    assert n == 2 and list(idx[:2]) == [0, 2] and list(cls[:2]) == [7, 1]
    np.testing.assert_allclose(bbox[0], [np.float32(0.2), np.float32(0.1), np.float32(0.8) - np.float64(np.float32(0.2)), np.float64(np.float32(0.5)) - np.float32(0.1)])
    assert (mu8[0] == 127).all() and (mu8[1] == 191).all()   # 255 - p/2*255 truncated (Detection.swift:84)


def test_mask_select(orc):
    rng = np.random.default_rng(5)
    m = rng.uniform(size=(4, 6, 28, 28)).astype(np.float32)
    det = np.zeros((4, 6), np.float32)
    det[:, 4] = [2, 5, 1, 0]
    out = orc.mask_select(m, [1, 1, 1, 0], det)
    np.testing.assert_array_equal(out[0], m[0, 2])
    np.testing.assert_array_equal(out[1], m[1, 5])
    np.testing.assert_array_equal(out[2], m[2, 1])
    assert not out[3].any()


def test_proposal_small_end_to_end(orc):
    # 6 anchors, pre-NMS limit 4: the two lowest scores never reach NMS
    anchors = np.array([[0.0, 0.0, 0.5, 0.5], [0.0, 0.01, 0.5, 0.51], [0.5, 0.5, 1.0, 1.0],
                        [0.2, 0.2, 0.4, 0.4], [0.0, 0.5, 0.5, 1.0], [0.5, 0.0, 1.0, 0.5]], np.float32)
    probs = np.zeros((6, 2), np.float32)
    probs[:, 1] = [0.9, 0.8, 0.7, 0.1, 0.2, 0.6]
    probs[:, 0] = 1 - probs[:, 1]
    deltas = np.zeros((6, 4), np.float32)
    rois, keep, cnt = orc.proposal(probs, deltas, anchors, pre_nms=4, max_proposals=5)
    assert cnt == 3
    np.testing.assert_array_equal(keep, [0, 2, 5, -1, -1])
    np.testing.assert_array_equal(rois[:3], anchors[[0, 2, 5]])
    assert not rois[3:].any()


def test_letterbox_known_answers(orc, pkg):
    import ctypes as C
    rng = np.random.default_rng(3)
    # same size: identity
    img = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    np.testing.assert_array_equal(orc.letterbox(img, 64, 64), img)
    # 2:1 landscape into a square: scale = dst_w / src_w, rows padded evenly above and below (DetectionRenderer.swift:63-75)
    img = np.full((32, 64, 3), 200, np.uint8)
    out = orc.letterbox(img, 128, 128)
    assert (out[:32] == 0).all() and (out[96:] == 0).all() and (out[32:96] == 200).all()
    g = (C.c_double * 5)()
    assert pkg.lib().mrcnn_letterbox_geometry(32, 64, 128, 128, g) == 0
    assert list(g) == [2.0, 128.0, 64.0, 0.0, 32.0]
    # portrait: columns padded
    out = orc.letterbox(np.full((64, 32, 3), 9, np.uint8), 128, 128)
    assert (out[:, :32] == 0).all() and (out[:, 96:] == 0).all() and (out[:, 32:96] == 9).all()
    # exact 2x up-sampling of a horizontal ramp stays monotone and inside the source range
    ramp = np.tile(np.arange(0, 64, dtype=np.uint8)[None, :, None] * 4, (64, 1, 3))
    up = orc.letterbox(ramp, 128, 128)
    assert (np.diff(up[5, :, 0].astype(int)) >= 0).all() and up.max() <= ramp.max()
    # boxes map back: the full padded frame's image area becomes [0,1]
    boxes = np.array([[0.25, 0.0, 0.75, 1.0, 3.0, 0.9]], np.float32)
    back = np.zeros_like(boxes)
    assert pkg.lib().mrcnn_unletterbox_boxes(32, 64, 128, 128, pkg._cabi.ptr(boxes), 1, 6, pkg._cabi.ptr(back)) == 0
    np.testing.assert_allclose(back[0], [0, 0, 1, 1, 3, 0.9], atol=1e-6)


def test_internal_fp16_roialign_is_bounded_against_the_fp32_layer(orc):
    """The fused pipeline keeps feature maps and pooled features in fp16 (NHWC); the oracle variant that checks that
    kernel (orc_pyramid_roialign_nhwc_f16) is the SAME fp32 arithmetic on fp16 taps with an fp16 result -- not a
    restatement of anything in the reference, which samples fp32 maps (PyramidROIAlignLayer.swift:186-242).  This test
    bounds the two roundings it adds against the fp32 layer on the same fp32 maps: on fp16-representable maps only
    the result rounding remains (half an fp16 ulp: 2^-11 relative), on arbitrary fp32 maps each tap also moves by
    2^-11 relative, and a bilinear sample is a convex combination of its taps."""
    import maskrcnn_b200 as m
    maps = m.synth.feature_maps(3, 256, 256, channels=16)                       # fp32 CHW, N(0, 1)
    rois = m.synth.random_rois(300, 4, min_px=6, max_px=250, image_size=256, n_pad=5)
    for pool in (7, 14):
        want, lv = orc.pyramid_roialign(rois, maps, pool, 256, 256)                                  # the fp32 layer
        # (1) maps already fp16-representable: only the result is rounded
        m16 = [x.astype(np.float16) for x in maps]
        exact_in, _ = orc.pyramid_roialign(rois, [x.astype(np.float32) for x in m16], pool, 256, 256)
        got, lv2 = orc.pyramid_roialign_nhwc_f16(rois, [np.ascontiguousarray(x.transpose(1, 2, 0)) for x in m16], pool, 256, 256)
        got = got.astype(np.float32).transpose(0, 3, 1, 2)
        np.testing.assert_array_equal(lv, lv2)
        np.testing.assert_array_equal(got, exact_in.astype(np.float16).astype(np.float32))           # = round16(fp32 layer)
        assert (np.abs(got - exact_in) <= 2.0 ** -11 * np.abs(exact_in) + 2.0 ** -25).all()
        # (2) arbitrary fp32 maps: taps rounded to fp16 as well; |sample error| <= 2^-11 max|tap| (+ result rounding)
        bound = 2.0 ** -11 * max(float(np.abs(x).max()) for x in maps) + 2.0 ** -11 * np.abs(want) + 2.0 ** -24
        assert (np.abs(got - want) <= bound).all()
        assert np.abs(got - want).max() > 0                                                           # the roundings are real
