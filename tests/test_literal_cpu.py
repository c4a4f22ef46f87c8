"""Two independent restatements of the reference must agree: oracle/oracle.c ("intended" mode, fused C loops: the checker
of the CUDA path) against oracle/literal.py (a statement-level numpy model of the Swift sources that keeps the
Accelerate calls, the Float-typed indices, the CGRect NMS loop and the ROIAlign group/batch bookkeeping, including the
behaviour SURVEY.md Appendix A lists as bugs).  Bit-exact wherever the reference is well defined; where it is not, the
tests pin down exactly what `intended` changes (Q4 trap, Q5 unwritten blocks, Q11/Q12 order freedom)."""
import numpy as np
import pytest

from oracle import literal as lit


@pytest.fixture(scope="module")
def m():
    import maskrcnn_b200
    return maskrcnn_b200


def _rpn_case(m, size, seed):
    anchors = m.synth.generate_anchors(size, size)
    probs, deltas = m.synth.rpn_outputs(anchors, seed, image_size=size)
    return anchors, probs, deltas


@pytest.mark.parametrize("size,pre,mx,seed", [(128, 600, 100, 0), (128, 600, 40, 1), (256, 1500, 200, 2)])
def test_proposal_literal_equals_oracle(orc, m, size, pre, mx, seed):
    anchors, probs, deltas = _rpn_case(m, size, seed)
    rois, keep, cnt, sorted_boxes = orc.proposal(probs, deltas, anchors, pre_nms=pre, max_proposals=mx, return_sorted_boxes=True)
    out, result_indices, order = lit.proposal_evaluate(probs, deltas, anchors, pre_nms=pre, max_proposals=mx, fix_q4=True)
    assert cnt == len(result_indices)
    np.testing.assert_array_equal(order[result_indices].astype(np.int32), keep[:cnt])   # anchor index of every kept roi
    np.testing.assert_array_equal(out, rois)                                            # boxes + zero padding, bit for bit
    if cnt == mx:
        # enough survivors: the reference's own index range (4x too long, Q4) is never walked past n -> same result
        out4, ri4, _ = lit.proposal_evaluate(probs, deltas, anchors, pre_nms=pre, max_proposals=mx, fix_q4=False)
        assert ri4 == result_indices
        np.testing.assert_array_equal(out4, rois)


def test_proposal_q4_trap_is_the_only_divergence(orc, m):
    """Fewer survivors than maxProposals: the reference indexes past its box array and traps (Q4); `intended` stops at n."""
    anchors, probs, deltas = _rpn_case(m, 128, 3)
    deltas = np.zeros_like(deltas)                      # undecoded anchors overlap heavily -> few survivors
    with pytest.raises(lit.SwiftTrap):
        lit.proposal_evaluate(probs, deltas, anchors, pre_nms=64, max_proposals=100, fix_q4=False)
    out, ri, order = lit.proposal_evaluate(probs, deltas, anchors, pre_nms=64, max_proposals=100, fix_q4=True)
    rois, keep, cnt = orc.proposal(probs, deltas, anchors, pre_nms=64, max_proposals=100)
    assert 0 < cnt < 100 and cnt == len(ri)
    np.testing.assert_array_equal(out, rois)
    np.testing.assert_array_equal(order[ri].astype(np.int32), keep[:cnt])


def test_proposal_tie_order_and_float_indices(orc, m):
    """Duplicated scores: stable descending order (lower anchor first); indices travel as Float through
    broadcastedIndices / vDSP_vindex exactly (Q2)."""
    anchors, probs, deltas = _rpn_case(m, 128, 4)
    probs = probs.copy()
    probs[:, 1] = np.round(probs[:, 1] * 16) / 16       # heavy ties
    probs[:, 0] = 1 - probs[:, 1]
    rois, keep, cnt = orc.proposal(probs, deltas, anchors, pre_nms=600, max_proposals=100)
    out, ri, order = lit.proposal_evaluate(probs, deltas, anchors, pre_nms=600, max_proposals=100, fix_q4=True)
    np.testing.assert_array_equal(out, rois)
    np.testing.assert_array_equal(order[ri].astype(np.int32), keep[:cnt])
    idx = np.array([0, 1, 4194303], np.float32)          # largest anchor index whose 4*i+3 is exact in Float
    b = lit.broadcasted_indices(idx, 4)
    np.testing.assert_array_equal(b.astype(np.int64), np.array([0, 1, 2, 3, 4, 5, 6, 7, 16777212, 16777213, 16777214, 16777215]))


def _det_case(m, seed, r=200, cluster=True):
    pr, bb = m.synth.classifier_outputs(r, seed)
    rois = m.synth.random_rois(r, seed, min_px=10, max_px=100, image_size=128)
    return rois, pr, bb


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_classifier_and_detection_literal_equal_oracle(orc, m, seed):
    rois, pr, bb = _det_case(m, seed)
    cls = lit.classifier_process_output(pr, bb)
    np.testing.assert_array_equal(cls, orc.classifier_select(pr, bb))
    cls[:, :4] *= np.float32(0.5)
    rng = np.random.default_rng(seed)
    rois[40:110] = np.clip(rois[40] + rng.uniform(-0.02, 0.02, (70, 4)).astype(np.float32), 0, 1)   # one crowded class
    cls[40:110, 4] = 7.0
    cls[40:110, 5] = rng.uniform(0.7, 1.0, 70).astype(np.float32)
    det, keep, cnt = orc.detection(rois, cls)
    out, roi_idx = lit.detection_evaluate(rois, cls)
    assert cnt == len(roi_idx) and cnt > 5
    np.testing.assert_array_equal(out, det)
    np.testing.assert_array_equal(np.array(roi_idx, np.int32), keep[:cnt])
    # Q11: any iteration order of the class Set gives the same rows when the kept scores are distinct
    kept_scores = det[:cnt, 5]
    if len(set(kept_scores.tolist())) == cnt:
        shuffled, _ = lit.detection_evaluate(rois, cls, class_order=lambda s: sorted(s, reverse=True))
        np.testing.assert_array_equal(shuffled, det)


def test_detection_order_freedom_only_permutes_equal_scores(orc, m):
    """Q11/Q12: with tied scores the reference's output order is not defined (Set order, unstable sort); every order it
    could produce differs from `intended` only by a permutation inside runs of equal score -- as long as the cut at
    maxDetections does not fall inside such a run."""
    rois, pr, bb = _det_case(m, 5)
    cls = orc.classifier_select(pr, bb)
    cls[:, :4] *= np.float32(0.25)
    cls[:, 5] = np.where(cls[:, 5] >= 0.7, np.float32(0.75), cls[:, 5])         # all survivors tie
    det, keep, cnt = orc.detection(rois, cls)
    assert 1 < cnt < 100
    for kw in (dict(class_order=lambda s: sorted(s, reverse=True)), dict(final_sort_reverse_ties=True)):
        out, idx = lit.detection_evaluate(rois, cls, **kw)
        assert len(idx) == cnt
        assert sorted(idx) == sorted(keep[:cnt].tolist())
        assert sorted(map(tuple, out[:cnt].tolist())) == sorted(map(tuple, det[:cnt].tolist()))
        np.testing.assert_array_equal(out[cnt:], 0)


def test_detection_threshold_edges(orc):
    """score >= Float(0.7) is kept (vDSP_vthres, Q8), background never is, classId > 0 is read from the Float array."""
    t = np.float32(0.7)
    rois = np.array([[0.1, 0.1, 0.3, 0.3], [0.5, 0.5, 0.8, 0.8], [0.2, 0.6, 0.4, 0.9], [0.6, 0.1, 0.9, 0.3]], np.float32)
    cls = np.zeros((4, 6), np.float32)
    cls[:, 4] = [3, 0, 5, 9]
    cls[:, 5] = [t, 0.99, np.nextafter(t, np.float32(0)), 0.8]
    out, idx = lit.detection_evaluate(rois, cls)
    det, keep, cnt = orc.detection(rois, cls)
    assert idx == [3, 0] and cnt == 2
    np.testing.assert_array_equal(out, det)


def _roialign_case(m, seed, n=64, n_pad=6, size=128):
    maps = m.synth.feature_maps(seed, size, size, channels=8)
    rois = m.synth.random_rois(n, seed, min_px=4, max_px=120, image_size=size, n_pad=n_pad)
    return rois, maps


@pytest.mark.parametrize("pool,stride", [(7, 4), (14, 6)])
def test_roialign_written_elements_equal_oracle_and_q5_is_the_rest(orc, m, pool, stride):
    rois, maps = _roialign_case(m, 7)
    rois[10] = [0.25, 0.25, 0.25, 0.75]                 # zero height -> a padding item in the middle
    rois[11] = [0.5, 0.5, 0.5, 0.5]
    if stride == 6:
        rois = np.concatenate([rois, np.ones((rois.shape[0], 2), np.float32)], axis=1)
    intended, levels = orc.pyramid_roialign(rois, maps, pool, 1024, 1024)
    out, written = lit.pyramid_roialign_evaluate(rois, maps, pool, 1024.0, 1024.0)
    # 1. level choice and validity of every roi
    items = lit.rois_to_input_items(rois, 224.0, 1024.0, 1024.0)
    lit_levels = np.array([c[0] + 2 if c is not None else -1 for _, c in items], np.int32)
    np.testing.assert_array_equal(lit_levels, levels)
    # 2. everything the reference writes has the intended value, bit for bit
    np.testing.assert_array_equal(out[written], intended[written])
    assert written.any()
    # 3. what it does not write: (a) the run that is still open when the rois end (Q5) -- here the trailing padding
    #    rows -- and (b) all but the first `count` floats of a padding run in the middle (copyOutput clears `count`
    #    floats, not `count` blocks).  `intended` writes zeros there.
    blk = written.reshape(written.shape[0], -1)
    full = blk.all(axis=1)
    n = rois.shape[0]
    assert not blk[n - 6:].any()                                        # (a) trailing padding run: never written
    assert blk[10, :2].all() and not blk[10, 2:].any() and not blk[11].any()       # (b) 2-roi padding run -> 2 floats
    assert (intended[~written] == 0).all()                              # here every unwritten element belongs to a padding roi
    np.testing.assert_array_equal(intended[10:12], 0)
    np.testing.assert_array_equal(intended[n - 6:], 0)
    # every roi before the last open run, other than the padding pair, is written in full
    last_group_start = _last_open_run_start(items)
    expect_full = np.array([i < last_group_start and i not in (10, 11) for i in range(n)])
    np.testing.assert_array_equal(full, expect_full)


def _last_open_run_start(items):
    """Index where the run that groupInputItemsByContent leaves open begins (same content kind / map up to the end,
    cut at multiples of 64 for region runs)."""
    key = lambda c: "p" if c is None else c[0]
    n = len(items)
    i = n - 1
    while i > 0 and key(items[i - 1][1]) == key(items[n - 1][1]):
        i -= 1
    run = n - i
    if items[n - 1][1] is not None:
        i += (run // 64) * 64
    return i


def test_roialign_full_groups_of_64_are_flushed(orc, m):
    """A run of exactly 64 same-level rois is closed inside the loop (:458-460): with 64 rois of one level at the end the
    reference writes everything, and the result equals `intended` in full."""
    maps = m.synth.feature_maps(3, 128, 128, channels=4)
    rng = np.random.default_rng(3)
    y = rng.uniform(0.0, 0.9, 64).astype(np.float32)
    x = rng.uniform(0.0, 0.9, 64).astype(np.float32)
    rois = np.stack([y, x, y + np.float32(0.05), x + np.float32(0.05)], axis=1).astype(np.float32)     # all level 2
    intended, levels = orc.pyramid_roialign(rois, maps, 7, 1024, 1024)
    assert (levels == 2).all()
    out, written = lit.pyramid_roialign_evaluate(rois, maps, 7, 1024.0, 1024.0)
    assert written.all()
    np.testing.assert_array_equal(out, intended)
    out63, written63 = lit.pyramid_roialign_evaluate(rois[:63], maps, 7, 1024.0, 1024.0)
    assert not written63.any()                          # one open group, never closed: the early return of :103-106


def test_roialign_half_levels_round_away_from_zero(orc):
    ratio = 224.0 / 1024.0
    rows = []
    for lv in (2.5, 3.5, 4.5):
        side = ratio * 2.0 ** (lv - 4.0)
        rows.append([0.0, 0.0, side, side])
    rois = np.array(rows, np.float32)
    items = lit.rois_to_input_items(rois, 224.0, 1024.0, 1024.0)
    np.testing.assert_array_equal(np.array([c[0] + 2 for _, c in items], np.int32), orc.roi_levels(rois))


def test_mask_layer_and_decode_literal_equal_oracle(orc):
    rng = np.random.default_rng(11)
    d, ncls, s = 20, 81, 8
    cnt = 13
    pooled = rng.normal(size=(d, 4, 4, 4)).astype(np.float32)
    pooled[cnt:] = 0                                                    # ROIAlign's zero blocks for padding detections
    det = np.zeros((d, 6), np.float32)
    det[:cnt, 4] = rng.integers(1, ncls, cnt)
    det[:cnt, 5] = rng.uniform(0.6, 1.0, cnt)
    det[:cnt, :4] = np.sort(rng.uniform(size=(cnt, 4)).astype(np.float32), axis=1)[:, [0, 1, 2, 3]]
    masks_all = rng.uniform(size=(d, ncls, s, s))                       # the Mask model emits Double
    out = lit.mask_evaluate(pooled, det, lambda i: masks_all[i])
    valid = (np.arange(d) < cnt).astype(np.int32)
    np.testing.assert_array_equal(out, orc.mask_select(masks_all.astype(np.float32), valid, det))
    # Q9 literally: a valid block holding one exact zero is dropped by the reference, the later blocks move up
    pooled2 = pooled.copy()
    pooled2[3, 0, 0, 0] = 0
    out2 = lit.mask_evaluate(pooled2, det, lambda i: masks_all[i])
    assert np.isnan(out2[3]).all()                                      # never written (stale memory) ...
    np.testing.assert_array_equal(out2[cnt - 1:], 0)                    # ... and the zeroed tail starts one row early
    # public Detection decoding
    big = rng.uniform(size=(d, 28, 28)).astype(np.float32)
    n, idx, bbox, dc, score, mu8 = orc.detections_decode(det, big)
    ref = lit.detections_from_feature_value(det, big.astype(np.float64))
    assert n == len(ref) and n > 0
    for k, (i, box, c, sc, mask) in enumerate(ref):
        assert i == idx[k] and c == dc[k] and sc == score[k]
        np.testing.assert_array_equal(np.array(box), bbox[k])
        np.testing.assert_array_equal(mask, mu8[k])


def test_literal_model_reproduces_the_committed_golden_fixtures():
    """tests/golden/*.npz were generated with oracle.c (make_golden.py) and are what the CUDA path is compared with on the
    GPU; the literal model of the Swift sources must give the same arrays, so the fixtures rest on both restatements."""
    import os
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(G, "proposal.npz"))
    for pk, rk, kk, ck in (("probs", "rois", "keep", "count"), ("tie_probs", "tie_rois", "tie_keep", "tie_count")):
        out, ri, order = lit.proposal_evaluate(g[pk], g["deltas"], g["anchors"], pre_nms=600, max_proposals=100, fix_q4=True)
        cnt = int(g[ck])
        assert len(ri) == cnt
        np.testing.assert_array_equal(out, g[rk])
        np.testing.assert_array_equal(order[ri].astype(np.int32), g[kk][:cnt])
    g = np.load(os.path.join(G, "roialign.npz"))
    maps = [g[f"maps{i}"] for i in range(4)]
    for pool, key, rois in ((7, "pooled7", g["rois"]), (14, "pooled14", np.concatenate([g["rois"], np.zeros((64, 2), np.float32)], axis=1))):
        out, written = lit.pyramid_roialign_evaluate(rois, maps, pool, 1024.0, 1024.0)
        np.testing.assert_array_equal(out[written], g[key][written])
        assert (g[key][~written] == 0).all() and written.mean() > 0.8
    items = lit.rois_to_input_items(g["rois"], 224.0, 1024.0, 1024.0)
    np.testing.assert_array_equal(np.array([c[0] + 2 if c is not None else -1 for _, c in items], np.int32), g["levels"])
    g = np.load(os.path.join(G, "detection.npz"))
    out, idx = lit.detection_evaluate(g["rois"], g["cls"])
    cnt = int(g["count"])
    np.testing.assert_array_equal(out, g["det"])
    np.testing.assert_array_equal(np.array(idx, np.int32), g["keep"][:cnt])
    m = np.load(os.path.join(G, "mask_decode.npz"))
    sel = lit.mask_evaluate(np.where(m["valid"][:, None, None, None] > 0, np.float32(1), np.float32(0)) * np.ones((100, 1, 2, 2), np.float32),
                            g["det"], lambda i: m["masks_all"][i])
    np.testing.assert_array_equal(sel, m["selected"])
    ref = lit.detections_from_feature_value(g["det"], m["masks"].astype(np.float64))
    assert len(ref) == int(m["n"])
    for k, (i, box, c, sc, mask) in enumerate(ref):
        assert i == m["index"][k] and c == m["classes"][k] and sc == m["score"][k]
        np.testing.assert_array_equal(np.array(box), m["bbox"][k])
        np.testing.assert_array_equal(mask, m["mask_u8"][k])


@pytest.mark.parametrize("size,pre,mx", [(1024, 6000, 1000), (512, 6000, 300)])
def test_proposal_at_the_baseline_sizes_with_the_reference_index_range(orc, m, size, pre, mx):
    """BASELINE.json configs A (261 888 anchors, 6000 -> 1000) and S (65 472 anchors, 6000 -> 300) at their exact sizes, the
    literal model run with the reference's own 4x-too-long NMS index range (enough boxes survive, so it never traps)."""
    anchors = m.synth.generate_anchors(size, size)
    assert anchors.shape[0] == {1024: 261888, 512: 65472}[size]
    probs, deltas = m.synth.rpn_outputs(anchors, 0, image_size=size)
    rois, keep, cnt = orc.proposal(probs, deltas, anchors, pre_nms=pre, max_proposals=mx)
    out, ri, order = lit.proposal_evaluate(probs, deltas, anchors, pre_nms=pre, max_proposals=mx, fix_q4=False)
    assert cnt == mx == len(ri)
    np.testing.assert_array_equal(out, rois)
    np.testing.assert_array_equal(order[ri].astype(np.int32), keep)
    # and the layers behind it on those rois: level rule + pooled values of a few rois on full-size maps, DetectionLayer
    maps = m.synth.feature_maps(0, size, size, channels=4)
    pick = rois[:: mx // 25]
    pooled, lv = orc.pyramid_roialign(pick, maps, 7, size, size)
    items = lit.rois_to_input_items(pick, 224.0, float(size), float(size))
    np.testing.assert_array_equal(np.array([c[0] + 2 for _, c in items], np.int32), lv)
    for i, (_, (mi, box)) in enumerate(items):
        np.testing.assert_array_equal(lit.crop_and_resize_bilinear(maps[mi], box, 7), pooled[i])
    pr, bb = m.synth.classifier_outputs(mx, 0)
    cls = orc.classifier_select(pr, bb)
    det, dkeep, dcnt = orc.detection(rois, cls)
    dout, didx = lit.detection_evaluate(rois, cls)
    assert dcnt == len(didx) and dcnt > 0
    np.testing.assert_array_equal(dout, det)
    np.testing.assert_array_equal(np.array(didx, np.int32), dkeep[:dcnt])


def test_nms_on_inverted_boxes_follows_cgrect_standardisation(orc):
    """CGRect.width / height / minX / maxX are those of the standardised rectangle (Utils.swift:194-246 builds
    CGRect(x: x1, y: y1, width: x2 - x1, height: y2 - y1) and reads them): a box given with x2 < x1 or y2 < y1 -- only
    possible for caller-supplied anchors / rois, the layers' own decode keeps the order -- is selectable and overlaps
    like its mirror image.  Both restatements agree, IoU values and keep lists, bit for bit."""
    rng = np.random.default_rng(3)
    n = 400
    c = rng.uniform(0.2, 0.8, (n, 2)); e = rng.uniform(0.02, 0.2, (n, 2))
    boxes = np.concatenate([c - e, c + e], 1).astype(np.float32)
    flip_y, flip_x = rng.random(n) < 0.3, rng.random(n) < 0.3
    boxes[flip_y] = boxes[flip_y][:, [2, 1, 0, 3]]
    boxes[flip_x] = boxes[flip_x][:, [0, 3, 2, 1]]
    boxes[7] = [0.3, 0.3, 0.3, 0.5]                                # zero height: never selectable
    for i, j in rng.integers(0, n, (300, 2)):
        assert orc.iou(boxes[i], boxes[j]) == float(lit.IOU(lit.CGRect(boxes[i]), lit.CGRect(boxes[j])))
    a = boxes[3].copy(); a[[0, 2]] = a[[2, 0]]
    assert orc.iou(a, boxes[3]) == 1.0 and orc.iou(boxes[3], a) == 1.0          # a box and its mirror image coincide
    for thr in (0.3, 0.7):
        keep = orc.nms(boxes, np.arange(n), thr, n)
        assert keep.tolist() == lit.non_max_supression(boxes.reshape(-1), list(range(n)), thr, n)
        assert 7 not in keep and (flip_y | flip_x)[keep].any()
