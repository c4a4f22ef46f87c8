"""GPU tests: the CUDA path through the C ABI against the committed golden fixtures (tests/golden/*.npz)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_proposal_golden_gpu(pkg):
    g = np.load(os.path.join(G, "proposal.npz"))
    c = pkg.Context(image_h=128, image_w=128, pre_nms_max_proposals=600, max_proposals=100)
    c.set_anchors(g["anchors"])
    layer = pkg.ProposalLayer({"preNMSMaxProposals": 600, "maxProposals": 100}, context=c)
    for pk, rk, kk, ck in (("probs", "rois", "keep", "count"), ("tie_probs", "tie_rois", "tie_keep", "tie_count")):
        rois = np.full((1, 100, 4), 5.0, np.float32); keep = np.zeros((1, 100), np.int32); cnt = np.zeros(1, np.int32)
        layer.evaluate([g[pk][None], g["deltas"][None]], [rois], keep, cnt)
        assert cnt[0] == int(g[ck])
        np.testing.assert_array_equal(keep[0], g[kk])
        np.testing.assert_array_equal(rois[0], g[rk])
    c.close()


def test_roialign_golden_gpu(pkg, ctx):
    g = np.load(os.path.join(G, "roialign.npz"))
    maps = [g[f"maps{i}"][None] for i in range(4)]
    out = np.full((1, 64, 8, 7, 7), 3.0, np.float32); lv = np.zeros((1, 64), np.int32)
    pkg.PyramidROIAlignLayer({"poolSize": 7}, context=ctx).evaluate([g["rois"][None]] + maps, [out], level=lv)
    np.testing.assert_array_equal(lv[0], g["levels"])
    np.testing.assert_array_equal(out[0], g["pooled7"])
    r6 = np.concatenate([g["rois"], np.zeros((64, 2), np.float32)], axis=1)
    out14 = np.zeros((1, 64, 8, 14, 14), np.float32)
    pkg.PyramidROIAlignLayer({"poolSize": 14}, context=ctx).evaluate([r6[None]] + maps, [out14])
    np.testing.assert_array_equal(out14[0], g["pooled14"])


def test_detection_and_decode_golden_gpu(pkg, ctx):
    g = np.load(os.path.join(G, "detection.npz"))
    cls = np.zeros((1, 200, 6), np.float32)
    pkg.TimeDistributedClassifierLayer(context=ctx).select(g["probs"][None], g["bbox"][None], cls)
    assert cls[0, 3, 4] == 5
    det = np.ones((1, 100, 6), np.float32); keep = np.zeros((1, 100), np.int32); cnt = np.zeros(1, np.int32)
    pkg.DetectionLayer(context=ctx).evaluate([g["rois"][None], g["cls"][None]], [det], keep, cnt)
    assert cnt[0] == int(g["count"])
    np.testing.assert_array_equal(keep[0], g["keep"])
    np.testing.assert_array_equal(det[0], g["det"])
    m = np.load(os.path.join(G, "mask_decode.npz"))
    dets = pkg.Detection.detectionsFromFeatureValue(det[0], m["masks"], context=ctx)
    n = int(m["n"])
    assert len(dets) == n
    for i, d in enumerate(dets):
        assert d.index == m["index"][i] and d.classId == m["classes"][i] and d.score == m["score"][i]
        np.testing.assert_array_equal(np.array(d.boundingBox), m["bbox"][i])
        np.testing.assert_array_equal(d.mask.reshape(-1), m["mask_u8"][i])
