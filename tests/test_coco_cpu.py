"""COCO instances reader (COCO.swift) and the evaluation loop (EvaluateCommand.swift:155-198) as host logic: a fake
model stands in for libmaskrcnn_cuda.so here; tests/test_zz_coco_eval_gpu.py runs `check_loop_against_direct_calls` on the
real one."""
import json

import numpy as np
import pytest


def _instances(tmp_path, images, annotations, info=None):
    doc = {"info": info if info is not None else {"description": "COCO 2017 Dataset", "url": "http://cocodataset.org", "version": "1.0",
                                                   "year": 2017, "contributor": "COCO Consortium", "date_created": "2017/09/01"},
           "licenses": [], "images": images, "annotations": annotations, "categories": [{"id": 1, "name": "person"}]}
    p = tmp_path / "instances_val2017.json"
    p.write_text(json.dumps(doc))
    return str(p)


def _image(i, w=64, h=48):
    return {"id": i, "file_name": f"{i:012d}.png", "width": w, "height": h, "coco_url": "x", "license": 1}


def test_coco_reader_and_iterator(pkg, tmp_path):
    images = [_image(i) for i in (785, 139, 724, 285, 632, 872, 397)]
    anns = [{"id": 1, "image_id": 139, "category_id": 64, "bbox": [1, 2, 3, 4], "iscrowd": 0, "area": 12.0},
            {"id": 2, "image_id": 139, "category_id": 72, "bbox": [5.5, 6, 7, 8], "iscrowd": 0},
            {"id": 3, "image_id": 872, "category_id": 1, "bbox": [0, 0, 1, 1], "iscrowd": 1}]
    coco = pkg.coco.COCO(_instances(tmp_path, images, anns))
    assert coco.instances.info.year == 2017 and len(coco.instances.images) == 7
    assert coco.instances.images[0].fileName == "000000000785.png"          # .convertFromSnakeCase
    assert [a.id for a in coco.index[139]] == [1, 2] and coco.index[139][1].bbox == [5.5, 6.0, 7.0, 8.0]
    it = list(coco.makeImageIterator(limit=5, sortById=True))               # EvaluateCommand.swift:165
    assert [im.id for im, _ in it] == [139, 285, 397, 632, 724]
    assert [a.categoryId for a in it[0][1]] == [64, 72] and it[1][1] == []
    assert [im.id for im, _ in coco.makeImageIterator()] == [785, 139, 724, 285, 632, 872, 397]
    with pytest.raises(ValueError, match="limit"):
        coco.makeImageIterator(limit=8)


@pytest.mark.parametrize("damage", ["no_info_field", "no_file_name", "string_id", "not_json", "no_annotations"])
def test_coco_reader_rejects_what_jsondecoder_rejects(pkg, tmp_path, damage):
    images, anns, info = [_image(1)], [{"id": 1, "image_id": 1, "category_id": 1, "bbox": [0, 0, 1, 1]}], None
    if damage == "no_info_field":
        info = {"description": "d", "url": "u", "version": "1", "year": 2017}       # contributor missing
    if damage == "no_file_name":
        del images[0]["file_name"]
    if damage == "string_id":
        images[0]["id"] = "1"
    path = _instances(tmp_path, images, anns, info)
    if damage == "not_json":
        open(path, "w").write("{ images: ")
    if damage == "no_annotations":
        doc = json.load(open(path))
        del doc["annotations"]
        open(path, "w").write(json.dumps(doc))
    with pytest.raises(ValueError, match="COCO instances"):
        pkg.coco.COCO(path)


class _FakeModel:
    """Shape-compatible stand-in for MaskRCNN: letterbox by nearest sampling, detections derived from the image."""
    shape = (32, 32, 3)

    def __init__(self):
        self.seen = []

    def letterbox(self, image):
        h, w = image.shape[:2]
        self.seen.append((h, w, int(image[0, 0, 0]), int(image[0, 0, 2])))
        ys = (np.arange(32) * h // 32)
        xs = (np.arange(32) * w // 32)
        return np.ascontiguousarray(image[ys][:, xs])

    def prediction(self, image):
        assert image.shape == self.shape and image.dtype == np.uint8
        det = np.zeros((100, 6), np.float32)
        k = int(image[0, 0, 0]) % 5 + 1
        for i in range(k):
            det[i] = [0.1 * i, 0.05 * i, 0.1 * i + 0.2, 0.05 * i + 0.3, i + 1, 0.95 - 0.05 * i]
        det[k] = [0, 0, 0.5, 0.5, 9, np.float32(0.7)]                       # not > 0.7 as Double: dropped
        return {"detections": det, "mask": np.zeros((100, 28, 28), np.float32)}


def test_evaluate_dataset_loop_writes_the_reference_results_message(pkg, tmp_path):
    cv2 = pytest.importorskip("cv2")
    ids = (785, 139, 724, 285, 632, 872)
    imgs = tmp_path / "val2017"
    imgs.mkdir()
    for i in ids:
        rgb = np.zeros((48, 64, 3), np.uint8)
        rgb[..., 0] = i % 251                                                # R
        rgb[..., 2] = 7                                                      # B
        assert cv2.imwrite(str(imgs / f"{i:012d}.png"), rgb[:, :, ::-1])     # OpenCV writes BGR
    path = _instances(tmp_path, [_image(i) for i in ids], [])
    model = _FakeModel()
    timed = []
    blob = pkg.coco.evaluate_dataset(model, path, str(imgs), "coco-val", limit=5, on_image=lambda im, s: timed.append(im.id))
    order = [139, 285, 632, 724, 785]
    assert timed == order
    assert model.seen == [(48, 64, i % 251, 7) for i in order]               # decoded as RGB, in id order
    back = pkg.results_pb.decode_results(blob)
    assert [r["imageInfo"] for r in back] == [{"datasetId": "coco-val", "id": str(i), "width": 64, "height": 48} for i in order]
    for r, i in zip(back, order):
        k = (i % 251) % 5 + 1
        assert len(r["detections"]) == k
        d = r["detections"][-1]
        j = k - 1
        assert d["classId"] == k and d["classLabel"] == "test" and d["probability"] == float(np.float32(0.95 - 0.05 * j))
        assert d["boundingBox"]["x"] == float(np.float32(0.05 * j)) and d["boundingBox"]["y"] == float(np.float32(0.1 * j))
        assert d["boundingBox"]["width"] == float(np.float32(0.05 * j + 0.3)) - float(np.float32(0.05 * j))
    with pytest.raises(ValueError, match="cannot decode"):
        (imgs / f"{139:012d}.png").write_bytes(b"not a png")
        pkg.coco.evaluate_dataset(model, path, str(imgs), "coco-val", limit=1)


def check_loop_against_direct_calls(pkg, model, tmp_path, size=(48, 64)):
    """evaluate_dataset == per-image letterbox + prediction + the reference's detection conversion, in id order."""
    import cv2
    ids = (42, 7, 19)
    imgs = tmp_path / "images"
    imgs.mkdir()
    rng = np.random.default_rng(3)
    pixels = {}
    for i in ids:
        pixels[i] = rng.integers(0, 256, size + (3,), dtype=np.uint8)
        assert cv2.imwrite(str(imgs / f"{i:012d}.png"), pixels[i][:, :, ::-1])
    path = _instances(tmp_path, [_image(i, w=size[1], h=size[0]) for i in ids], [])
    blob = pkg.coco.evaluate_dataset(model, path, str(imgs), "ds", limit=3)
    want = []
    for i in sorted(ids):
        det = model.prediction(model.letterbox(pixels[i]))["detections"]
        want.append(pkg.results_pb.result_from_detections("ds", i, size[1], size[0], det))
    assert blob == pkg.results_pb.encode_results(want)
    return pkg.results_pb.decode_results(blob)


def test_loop_equals_direct_calls_with_the_fake_model(pkg, tmp_path):
    pytest.importorskip("cv2")
    back = check_loop_against_direct_calls(pkg, _FakeModel(), tmp_path)
    assert [r["imageInfo"]["id"] for r in back] == ["7", "19", "42"]
