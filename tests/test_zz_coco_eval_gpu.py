"""The evaluation loop (coco.evaluate_dataset: COCO reader -> letterbox -> MaskRCNN -> Results protobuf) on the real
library: a small synthetic-weight model (ResNet50, 256x256) on three PNG files of another size."""
import pytest

from test_coco_cpu import check_loop_against_direct_calls

pytestmark = pytest.mark.gpu


def test_evaluate_dataset_on_the_real_model_gpu(pkg, tmp_path):
    pytest.importorskip("cv2")
    cfg = pkg.MaskRCNNConfig()
    cfg.architecture, cfg.imageShape, cfg.preNMSMaxProposals, cfg.maxProposals, cfg.maxBatch = "resnet50", (256, 256, 3), 1000, 200, 2
    _, blobs = pkg.weights.synthetic_blobs(50)
    model = pkg.MaskRCNN(cfg, blobs=blobs, anchors=pkg.synth.generate_anchors(256, 256))
    try:
        back = check_loop_against_direct_calls(pkg, model, tmp_path, size=(180, 320))
        assert [r["imageInfo"] for r in back] == [{"datasetId": "ds", "id": str(i), "width": 320, "height": 180} for i in (7, 19, 42)]
        for r in back:
            for d in r["detections"]:
                assert d["probability"] > 0.7 and d["classId"] >= 1 and d["classLabel"] == "test"
    finally:
        model.close()


def test_anchors_generated_on_demand_gpu(pkg):
    """No anchors.bin configured: MaskRCNN generates the anchors for its input size (mrcnn_generate_anchors, the reference's
    own TODO) -- same predictions as with the anchors passed in."""
    import numpy as np
    cfg = pkg.MaskRCNNConfig()
    cfg.architecture, cfg.imageShape, cfg.preNMSMaxProposals, cfg.maxProposals, cfg.maxBatch = "resnet50", (256, 256, 3), 1000, 200, 2
    _, blobs = pkg.weights.synthetic_blobs(50)
    img = np.random.default_rng(20260).integers(0, 256, (256, 256, 3), dtype=np.uint8)
    a = pkg.MaskRCNN(cfg, blobs=blobs)
    b = pkg.MaskRCNN(cfg, blobs=blobs, anchors=pkg.synth.generate_anchors(256, 256))
    try:
        assert int(pkg.lib().mrcnn_num_anchors(a.ctx.handle)) == 16368
        oa, ob = a.prediction(img), b.prediction(img)
        np.testing.assert_array_equal(oa["detections"], ob["detections"])
        np.testing.assert_array_equal(oa["mask"], ob["mask"])
    finally:
        a.close()
        b.close()
