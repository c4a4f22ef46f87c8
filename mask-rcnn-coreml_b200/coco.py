"""COCO instances reader and the evaluation loop that feeds the reference's COCOEval task -- SURVEY.md section 8 f3.

`COCO` mirrors Sources/maskrcnn/COCO.swift (the Codable structs :79-107 decoded with `.convertFromSnakeCase`, the
image-id index :9-27, `makeImageIterator(limit:sortById:)` :60-77).  `evaluate_dataset` is the loop of
EvaluateCommand.swift:155-198: the first `limit` images by id, Vision's `.scaleFit` (here mrcnn_letterbox_eval),
one prediction per image, detections with Double(score) > 0.7 converted as :203-248 into a `Results` message
(results_pb), whose bytes Python/COCOEval/task.py:93-96 parses unchanged.  Boxes stay normalised to the letter-boxed
model frame, as in the reference (it never maps them back).

Host logic only; the predictions come from `MaskRCNN` (libmaskrcnn_cuda.so).  Not the reference's CLI: no argument
parsing, Docker or download steps.
"""
import json
import os
from collections import namedtuple

from . import results_pb

COCOInfo = namedtuple("COCOInfo", "description url version year contributor")
COCOImage = namedtuple("COCOImage", "id fileName width height")
COCOAnnotation = namedtuple("COCOAnnotation", "id imageId categoryId bbox")
COCOInstances = namedtuple("COCOInstances", "info images annotations")


def _need(d, key, kind, where):
    """JSONDecoder throws keyNotFound / typeMismatch; so do we."""
    if not isinstance(d, dict) or key not in d:
        raise ValueError(f"COCO instances: key '{key}' missing in {where}")
    v = d[key]
    if kind is int and (isinstance(v, bool) or not isinstance(v, int)):
        raise ValueError(f"COCO instances: '{key}' in {where} is not an integer")
    if kind is str and not isinstance(v, str):
        raise ValueError(f"COCO instances: '{key}' in {where} is not a string")
    return v


class COCO:
    """COCO.swift:3-77."""

    def __init__(self, url):
        with open(url, "rb") as f:
            try:
                doc = json.load(f)
            except json.JSONDecodeError as e:
                raise ValueError(f"COCO instances: not valid JSON ({e})") from None
        info = _need(doc, "info", None, "the document")
        self.instances = COCOInstances(
            COCOInfo(_need(info, "description", str, "info"), _need(info, "url", str, "info"), _need(info, "version", str, "info"),
                     _need(info, "year", int, "info"), _need(info, "contributor", str, "info")),
            [COCOImage(_need(i, "id", int, "an image"), _need(i, "file_name", str, "an image"), _need(i, "width", int, "an image"),
                       _need(i, "height", int, "an image")) for i in _need(doc, "images", None, "the document")],
            [COCOAnnotation(_need(a, "id", int, "an annotation"), _need(a, "image_id", int, "an annotation"),
                            _need(a, "category_id", int, "an annotation"), [float(x) for x in _need(a, "bbox", None, "an annotation")])
             for a in _need(doc, "annotations", None, "the document")])
        self._index = None

    @property
    def index(self):
        """annotationsByImageIds (COCO.swift:9-27), built on first use (`lazy var`, :49-51)."""
        if self._index is None:
            by_image = {}
            for a in self.instances.annotations:
                by_image.setdefault(a.imageId, []).append(a)
            self._index = by_image
        return self._index

    def makeImageIterator(self, limit=None, sortById=False):
        """COCO.swift:60-77: yields (COCOImage, [COCOAnnotation])."""
        images = list(self.instances.images)
        if sortById:
            images.sort(key=lambda i: i.id)
        if limit is not None:
            if limit > len(images) or limit < 0:
                raise ValueError(f"limit {limit} out of range: {len(images)} images (the reference's slice traps)")
            images = images[:limit]
        index = self.index
        return ((image, index.get(image.id, [])) for image in images)


def read_image_rgb(path):
    """(H,W,3) u8 RGB.  The reference decodes with CIImage(contentsOf:) (EvaluateCommand.swift:169); here OpenCV."""
    import cv2
    bgr = cv2.imread(path, cv2.IMREAD_COLOR)
    if bgr is None:
        raise ValueError(f"cannot decode image {path}")
    return bgr[:, :, ::-1].copy()


def evaluate_dataset(model, instances_url, images_directory, dataset_id, limit=5, read_image=read_image_rgb, on_image=None):
    """EvaluateCommand.swift:155-198 -> serialized `Results` bytes.
    model: MaskRCNN (uses .letterbox and .prediction); on_image(image, seconds) replaces the reference's print (:193)."""
    import time
    coco = COCO(instances_url)
    results = []
    for image, _annotations in coco.makeImageIterator(limit=limit, sortById=True):      # :165
        start = time.perf_counter()
        rgb = read_image(os.path.join(images_directory, image.fileName))
        out = model.prediction(model.letterbox(rgb))                                    # .scaleFit + the model (:157,171)
        seconds = time.perf_counter() - start
        results.append(results_pb.result_from_detections(dataset_id, image.id, image.width, image.height, out["detections"]))
        if on_image is not None:
            on_image(image, seconds)
    return results_pb.encode_results(results)
