"""The reference's public surface, host side: MaskRCNNConfig, Detection and the MaskRCNN model
class (the Xcode-generated class over MaskRCNN.mlmodel: Example/Source/ViewController.swift:37),
forwarding to mrcnn_predict / mrcnn_detections_decode of libmaskrcnn_cuda.so.
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import check, lib, ptr
from .layers import Context


class MaskRCNNConfig:
    """MaskRCNNConfig.swift:10-18: a global singleton with the artefact URLs.  The Swift class has only the
    three URL properties; the layer parameters that Core ML bakes into the .mlmodel (ProposalLayer.swift:57-63,
    DetectionLayer.swift:55-61, PyramidROIAlignLayer.swift:45-46) live here too, with the Swift defaults."""
    defaultConfig = None

    def __init__(self):
        self.anchorsURL = None                    # MaskRCNNConfig.swift:15
        self.compiledClassifierModelURL = None    # :16
        self.compiledMaskModelURL = None          # :17
        self.modelURL = None                      # the MaskRCNN model bundle itself (ViewController.swift:37)
        self.architecture = "resnet101"           # README.md:87
        self.imageShape = (1024, 1024, 3)         # README.md:88
        self.numClasses = 81                      # README.md:89
        self.preNMSMaxProposals = 6000            # ProposalLayer.swift:59
        self.maxProposals = 1000                  # ProposalLayer.swift:61
        self.maxDetections = 100                  # DetectionLayer.swift:57
        self.maxBatch = 8
        self.preciseMasks = True                  # 2-term fp16 activations in the mask head (masks within 1e-4 of fp32); False: 1-term, 5e-4, ~8 % faster

    def context_overrides(self):
        return dict(image_h=self.imageShape[0], image_w=self.imageShape[1],
                    architecture={"resnet101": 101, "resnet50": 50}[self.architecture],
                    num_classes=self.numClasses, pre_nms_max_proposals=self.preNMSMaxProposals,
                    max_proposals=self.maxProposals, max_detections=self.maxDetections, max_batch=self.maxBatch, precise_masks=int(self.preciseMasks),
                    anchors_path=self.anchorsURL, main_model_path=self.modelURL,
                    classifier_model_path=self.compiledClassifierModelURL, mask_model_path=self.compiledMaskModelURL)


MaskRCNNConfig.defaultConfig = MaskRCNNConfig()


class Detection:
    """Detection.swift:15-21."""
    __slots__ = ("index", "boundingBox", "classId", "score", "mask")

    def __init__(self, index, boundingBox, classId, score, mask):
        self.index, self.boundingBox, self.classId, self.score, self.mask = index, boundingBox, classId, score, mask

    def __repr__(self):
        return f"Detection(index={self.index}, classId={self.classId}, score={self.score:.4f}, boundingBox={self.boundingBox})"

    @staticmethod
    def detectionsFromFeatureValue(detections, mask=None, context=None):
        """Detection.swift:23-62 (+ :64-99 for the 8-bit masks), computed on the device.
        detections (D,6) / mask (D,S,S) for one image -> [Detection]."""
        from .layers import default_context
        ctx = context or default_context()
        det = np.ascontiguousarray(detections, dtype=np.float32).reshape(1, -1, 6)
        d = det.shape[1]
        if d != ctx.cfg.max_detections:
            raise _cabi.MaskRCNNError(_cabi.EINVAL, "detections must have max_detections rows")
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.float32)
        s = 2 * ctx.cfg.pool_size_mask
        cnt = np.zeros(1, np.int32); idx = np.zeros((1, d), np.int32); bbox = np.zeros((1, d, 4), np.float64)
        cls = np.zeros((1, d), np.int32); score = np.zeros((1, d), np.float64)
        mu8 = np.zeros((1, d, s * s), np.uint8) if m is not None else None
        check(ctx.handle, lib().mrcnn_detections_decode(ctx.handle, 1, ptr(det), ptr(m), ptr(cnt), ptr(idx), ptr(bbox),
                                                        ptr(cls), ptr(score), ptr(mu8)))
        out = []
        for i in range(int(cnt[0])):
            out.append(Detection(int(idx[0, i]), tuple(bbox[0, i]), int(cls[0, i]), float(score[0, i]),
                                 mu8[0, i].reshape(s, s).copy() if mu8 is not None else None))
        return out


class MaskRCNN:
    """The model class: image (H,W,3) u8 -> outputs "detections" (100,6) and "mask" (100,28,28)
    (I/O names Conversion/task.py:70-72).  `configuration` defaults to MaskRCNNConfig.defaultConfig,
    like the custom layers of the reference read the singleton (ProposalLayer.swift:68)."""

    def __init__(self, configuration=None, device=None, blobs=None, anchors=None):
        cfg = configuration or MaskRCNNConfig.defaultConfig
        self.config = cfg
        ov = cfg.context_overrides()
        if device is not None:
            ov["device"] = device
        self.ctx = Context(**ov)
        if anchors is not None:
            self.ctx.set_anchors(anchors)
        elif cfg.anchorsURL is None:
            self.ctx.generate_anchors()            # no anchors.bin configured: generated for the model's input size
        if blobs is not None:                      # in-memory weights instead of files
            for which, blob in enumerate(blobs):
                self.ctx.set_weights(which, blob)
        c = self.ctx.cfg
        self.shape = (c.image_h, c.image_w, 3)
        self.D, self.S = c.max_detections, 2 * c.pool_size_mask

    def prediction_batch(self, images, detections=None, masks=None):
        """images [B,H,W,3] u8 (numpy / torch, host or cuda) -> (detections [B,D,6], masks [B,D,S,S]) f32.
        Output buffers may be passed in (same kinds); host outputs imply a stream synchronisation."""
        b = images.shape[0]
        if tuple(images.shape[1:]) != self.shape:
            raise _cabi.MaskRCNNError(_cabi.EINVAL, f"images must be [B,{self.shape[0]},{self.shape[1]},3] uint8")
        if detections is None:
            detections = np.empty((b, self.D, 6), np.float32)
        if masks is None:
            masks = np.empty((b, self.D, self.S, self.S), np.float32)
        check(self.ctx.handle, lib().mrcnn_predict(self.ctx.handle, b, *self._io(b, images, detections, masks)))
        return detections, masks

    def _io(self, b, images, detections, masks, total=None):
        """Checked pointers of the model's input / outputs (element type and size; see _cabi.ptr)."""
        t = b if total is None else total
        return (ptr(images, "uint8", b * self.shape[0] * self.shape[1] * 3), ptr(detections, "float32", t * self.D * 6),
                ptr(masks, "float32", t * self.D * self.S * self.S))

    # ---- streaming: two batches in flight (mrcnn_predict_submit / mrcnn_predict_wait) ----
    def submit(self, images, detections, masks, allgather=False):
        """Enqueues one batch and returns at once; the host->device copy overlaps the previous batch's compute.
        Host buffers should be pinned and must stay untouched until the matching wait() returns.  With
        allgather=True the outputs hold every rank's results (mrcnn_predict_allgather semantics)."""
        b = images.shape[0]
        if tuple(images.shape[1:]) != self.shape:
            raise _cabi.MaskRCNNError(_cabi.EINVAL, f"images must be [B,{self.shape[0]},{self.shape[1]},3] uint8")
        flags = 1 if allgather else 0          # MRCNN_SUBMIT_ALLGATHER
        total = b * max(int(getattr(self.ctx, "nranks", 1)), 1) if allgather else b
        check(self.ctx.handle, lib().mrcnn_predict_submit(self.ctx.handle, b, *self._io(b, images, detections, masks, total), flags))

    def wait(self):
        """Blocks until the oldest submitted batch is complete (outputs are in the buffers given to submit)."""
        check(self.ctx.handle, lib().mrcnn_predict_wait(self.ctx.handle))

    @property
    def in_flight(self):
        return int(lib().mrcnn_predict_in_flight(self.ctx.handle))

    def prediction_stream(self, batches):
        """Iterator of image batches -> iterator of (detections, masks), in order, keeping two batches in flight:
        the loop of EvaluateCommand.swift:166-194 with the copies off the critical path.
        Outputs land in PINNED host slots (three, reused) and are yielded as copies: with pageable buffers the
        device-to-host copy would block submit() until the batch is computed, and the overlap would be lost.  Input
        batches are used as given -- pin them (torch .pin_memory()) for an asynchronous upload.  Closing the generator
        early waits for the batches still in flight, so their buffers are not released under a running copy."""
        import torch
        slots, pending, n = [], [], 0
        try:
            for images in batches:
                b = images.shape[0]
                if len(pending) == 2:
                    self.wait()
                    _, det, msk = pending.pop(0)
                    yield det.numpy().copy(), msk.numpy().copy()
                k = n % 3
                n += 1
                if len(slots) <= k or slots[k][0].shape[0] != b:
                    slot = (torch.empty((b, self.D, 6), dtype=torch.float32).pin_memory(),
                            torch.empty((b, self.D, self.S, self.S), dtype=torch.float32).pin_memory())
                    if len(slots) <= k:
                        slots.append(slot)
                    else:
                        slots[k] = slot
                det, msk = slots[k]
                self.submit(images, det, msk)
                pending.append((images, det, msk))      # keeps the input alive until its wait
            while pending:
                self.wait()
                _, det, msk = pending.pop(0)
                yield det.numpy().copy(), msk.numpy().copy()
        finally:
            while self.in_flight:                       # generator closed early (or an error): drain what was submitted
                try:
                    self.wait()
                except _cabi.MaskRCNNError:
                    break

    def prediction(self, image):
        """One image -> {"detections": (D,6), "mask": (D,S,S)} like MaskRCNNOutput."""
        d, m = self.prediction_batch(np.ascontiguousarray(image)[None])
        return {"detections": d[0], "mask": m[0]}

    def predict(self, image):
        """image -> [Detection] (score > 0.7), the call pattern of ViewController.swift:163-187."""
        out = self.prediction(image)
        return Detection.detectionsFromFeatureValue(out["detections"], out["mask"], context=self.ctx)

    def letterbox(self, image, out=None):
        """Vision's .scaleFit in front of the model (EvaluateCommand.swift:157): (H,W,3) u8 of any size -> model-sized image."""
        h, w = image.shape[:2]
        if out is None:
            out = np.empty(self.shape, np.uint8)
        if image.ndim != 3 or image.shape[2] != 3:
            raise _cabi.MaskRCNNError(_cabi.EINVAL, "letterbox: image must be (H, W, 3) uint8")
        check(self.ctx.handle, lib().mrcnn_letterbox_eval(self.ctx.handle, ptr(image, "uint8", h * w * 3), h, w,
                                                          ptr(out, "uint8", self.shape[0] * self.shape[1] * 3)))
        return out

    def unletterbox(self, rows, src_h, src_w):
        """(n, 4|6) rows with boxes normalised to the model frame -> normalised to the source image."""
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        out = np.empty_like(rows)
        st = lib().mrcnn_unletterbox_boxes(src_h, src_w, self.shape[0], self.shape[1], ptr(rows), rows.shape[0], rows.shape[1], ptr(out))
        check(self.ctx.handle, st)
        return out

    def close(self):
        self.ctx.close()


def smoke_predict():
    """Tiny end-to-end MaskRCNN.predict (ResNet50, 256x256) used by __graft_entry__.smoke()."""
    from . import synth, weights
    cfg = MaskRCNNConfig()
    cfg.architecture, cfg.imageShape, cfg.preNMSMaxProposals, cfg.maxProposals, cfg.maxBatch = "resnet50", (256, 256, 3), 1000, 200, 2
    _, blobs = weights.synthetic_blobs(50)
    model = MaskRCNN(cfg, blobs=blobs, anchors=synth.generate_anchors(256, 256))
    rng = np.random.default_rng(20260)
    img = rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)
    out = model.prediction(img)
    dets = Detection.detectionsFromFeatureValue(out["detections"], out["mask"], context=model.ctx)
    assert np.isfinite(out["detections"]).all() and np.isfinite(out["mask"]).all()
    print(f"smoke: MaskRCNN.predict ok, {len(dets)} detections, stages:",
          ", ".join(f"{n} {ms:.2f} ms" for n, ms in model.ctx.stage_times()))
    model.close()
