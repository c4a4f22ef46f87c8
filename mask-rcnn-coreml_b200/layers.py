"""Host-side mirror of the reference's custom-layer plugin interface.

The reference implements Apple's ``MLCustomLayer`` protocol with five classes
(``init(parameters:)``, ``setWeightData``, ``outputShapes(forInputShapes:)``,
``evaluate(inputs:outputs:)``).  The classes below keep those names, argument
meanings, parameter keys and defaults, and forward ``evaluate`` to the C ABI of
libmaskrcnn_cuda.so.  Buffers may be numpy arrays (host) or torch tensors (host
or cuda); outputs are written in place like Core ML's pre-allocated MLMultiArrays.

Shapes follow the reference's 5-D ``[seq, batch, channel, height, width]``
convention only in ``output_shapes``; ``evaluate`` takes dense arrays with an
optional leading image-batch dimension.
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import MaskRCNNError, check, lib, mrcnn_config, ptr


class Context:
    """Owns one ``mrcnn_ctx`` (one CUDA device + stream).  Not thread-safe."""

    def __init__(self, config=None, **overrides):
        l = lib()
        cfg = mrcnn_config()
        l.mrcnn_config_default(C.byref(cfg))
        self._keep = []
        if config is not None:
            for k, v in config.items():
                overrides.setdefault(k, v)
        for k, v in overrides.items():
            if v is None:
                continue
            if k in ("bbox_std", "mean_rgb"):
                arr = getattr(cfg, k)
                for i, x in enumerate(v):
                    arr[i] = float(x)
            elif k.endswith("_path"):
                b = str(v).encode()
                self._keep.append(b)
                setattr(cfg, k, b)
            else:
                setattr(cfg, k, v)
        self.cfg = cfg
        h = C.c_void_p()
        st = l.mrcnn_create(C.byref(cfg), C.byref(h))
        if st != 0:
            msg = l.mrcnn_last_error(None)
            raise MaskRCNNError(st, msg.decode() if msg else "mrcnn_create failed")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            lib().mrcnn_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_anchors(self, anchors):
        a = np.ascontiguousarray(anchors, dtype=np.float32).reshape(-1, 4)
        check(self.handle, lib().mrcnn_set_anchors(self.handle, ptr(a), a.shape[0]))

    def generate_anchors(self):
        """Anchors for the configured image size, generated on demand instead of read from anchors.bin (the reference's own
        TODO, MaskRCNNConfig.swift:14); mrcnn_generate_anchors + mrcnn_set_anchors.  Returns the (N,4) array."""
        l = lib()
        n = l.mrcnn_anchor_count(self.cfg.image_h, self.cfg.image_w)
        if n < 0:
            raise MaskRCNNError(int(n), "mrcnn_anchor_count: bad image size")
        a = np.empty((n, 4), np.float32)
        st = l.mrcnn_generate_anchors(self.cfg.image_h, self.cfg.image_w, ptr(a), n)
        if st != 0:
            raise MaskRCNNError(st, "mrcnn_generate_anchors failed")
        self.set_anchors(a)
        return a

    def set_weights(self, which, blob):
        buf = bytes(blob)
        check(self.handle, lib().mrcnn_set_weights(self.handle, which, buf, len(buf)))

    def set_stream(self, cuda_stream_handle):
        check(self.handle, lib().mrcnn_set_stream(self.handle, C.c_void_p(cuda_stream_handle)))

    def synchronize(self):
        check(self.handle, lib().mrcnn_synchronize(self.handle))

    @property
    def launch_count(self):
        return int(lib().mrcnn_launch_count(self.handle))

    def profile_enable(self, on=True):
        check(self.handle, lib().mrcnn_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self):
        """{kernel class: (ms, launches, algorithmic work)} since the last read (synchronises)."""
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        cnt = (C.c_int64 * 16)()
        work = (C.c_double * 16)()
        n = lib().mrcnn_profile_read(self.handle, 16, names, ms, cnt, work)
        return {names[i].decode(): (float(ms[i]), int(cnt[i]), float(work[i])) for i in range(n)}

    def stage_times(self):
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        n = lib().mrcnn_last_stage_times(self.handle, 32, names, ms)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]


_default_ctx = None


def default_context():
    """Process-wide context, the analogue of the layers sharing MaskRCNNConfig.defaultConfig."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def _batch_of(x, base_ndim):
    return (x.shape[0], True) if x.ndim == base_ndim + 1 else (1, False)


class _Layer:
    def __init__(self, parameters=None, context=None):
        self.parameters = dict(parameters or {})
        self._ctx = context

    @property
    def ctx(self):
        return self._ctx if self._ctx is not None else default_context()

    def set_weight_data(self, weights):
        """setWeightData(_:) is a no-op in every reference layer (e.g. ProposalLayer.swift:93-95)."""


def _std_dev(params, default):
    # ProposalLayer.swift:70-80 / DetectionLayer.swift:67-77
    n = params.get("bboxStdDev_count")
    if isinstance(n, int):
        vals = [params.get(f"bboxStdDev_{i}") for i in range(n)]
        if all(isinstance(v, float) for v in vals):
            return [float(np.float32(v)) for v in vals]
    return list(default)


class ProposalLayer(_Layer):
    """ProposalLayer.swift:52-197."""

    def __init__(self, parameters=None, context=None):
        super().__init__(parameters, context)
        p = self.parameters
        self.bounding_box_refinement_standard_deviation = _std_dev(p, [0.1, 0.1, 0.2, 0.2])
        self.pre_nms_max_proposals = p["preNMSMaxProposals"] if isinstance(p.get("preNMSMaxProposals"), int) else 6000
        self.max_proposals = p["maxProposals"] if isinstance(p.get("maxProposals"), int) else 1000
        self.nms_iou_threshold = float(np.float32(p["nmsIOUThreshold"])) if isinstance(p.get("nmsIOUThreshold"), float) else float(np.float32(0.7))
        # the context holds the configuration; layers with non-default parameters need their own Context
        self._check_cfg()

    def _check_cfg(self):
        c = self.ctx.cfg
        if (c.pre_nms_max_proposals, c.max_proposals) != (self.pre_nms_max_proposals, self.max_proposals) or \
                abs(c.proposal_nms_iou - self.nms_iou_threshold) > 0:
            raise MaskRCNNError(_cabi.EINVAL, "ProposalLayer parameters differ from the context configuration; "
                                "create Context(pre_nms_max_proposals=..., max_proposals=..., proposal_nms_iou=...)")

    def output_shapes(self, for_input_shapes):
        out = list(for_input_shapes[1])      # ProposalLayer.swift:97-101
        out[0] = self.max_proposals
        return [out]

    def evaluate(self, inputs, outputs, keep_anchor=None, count=None):
        probs, deltas = inputs
        rois = outputs[0]
        batch, _ = _batch_of(probs, 2)
        n = probs.shape[-2]
        check(self.ctx.handle, lib().mrcnn_proposal_eval(self.ctx.handle, batch, n, ptr(probs), ptr(deltas),
                                                         ptr(rois), ptr(keep_anchor), ptr(count)))


class PyramidROIAlignLayer(_Layer):
    """PyramidROIAlignLayer.swift:40-183."""

    def __init__(self, parameters=None, context=None):
        super().__init__(parameters, context)
        p = self.parameters
        self.pool_size = p["poolSize"] if isinstance(p.get("poolSize"), int) else 7
        # Q15: the reference reads imageWidth/imageHeight `as? CGFloat` although the converter writes ints;
        # intended behaviour = always honour the configured size.
        self.image_size = (self.ctx.cfg.image_w, self.ctx.cfg.image_h)

    def output_shapes(self, for_input_shapes):
        rois, fmap = for_input_shapes[0], for_input_shapes[1]
        return [[rois[0], rois[1], fmap[2], self.pool_size, self.pool_size]]

    def evaluate(self, inputs, outputs, level=None):
        rois = inputs[0]
        fmaps = list(inputs[1:5])
        out = outputs[0]
        batch, has_b = _batch_of(rois, 2)
        r, stride = rois.shape[-2], rois.shape[-1]
        fshape = fmaps[0].shape[1:] if has_b else fmaps[0].shape
        c = fshape[0]
        hw = (C.c_int32 * 8)()
        for l, f in enumerate(fmaps):
            hw[2 * l], hw[2 * l + 1] = f.shape[-2], f.shape[-1]
        fp = (C.c_void_p * 4)(*[ptr(f) for f in fmaps])
        check(self.ctx.handle, lib().mrcnn_pyramid_roialign_eval(self.ctx.handle, batch, ptr(rois), stride, r, fp, hw,
                                                                 c, self.pool_size, ptr(out), ptr(level)))


class TimeDistributedClassifierLayer(_Layer):
    """TimeDistributedClassifierLayer.swift:14-92."""

    def output_shapes(self, for_input_shapes):
        s = for_input_shapes[0]
        return [[s[0], s[1], 1, 1, 6]]

    def evaluate(self, inputs, outputs):
        pooled = inputs[0]
        batch, _ = _batch_of(pooled, 4)
        r = pooled.shape[-4]
        check(self.ctx.handle, lib().mrcnn_classifier_eval(self.ctx.handle, batch, r, ptr(pooled), ptr(outputs[0])))

    def select(self, probabilities, bounding_boxes, out):
        """The post-processing half (:50-88) on explicit Classifier outputs."""
        batch, _ = _batch_of(probabilities, 2)
        r = probabilities.shape[-2]
        check(self.ctx.handle, lib().mrcnn_classifier_select(self.ctx.handle, batch, r, ptr(probabilities),
                                                             ptr(bounding_boxes), ptr(out)))


class DetectionLayer(_Layer):
    """DetectionLayer.swift:52-236."""

    def __init__(self, parameters=None, context=None):
        super().__init__(parameters, context)
        p = self.parameters
        self.bounding_box_refinement_standard_deviation = _std_dev(p, [0.1, 0.1, 0.2, 0.2])
        self.max_detections = p["maxDetections"] if isinstance(p.get("maxDetections"), int) else 100
        self.low_confidence_score_threshold = float(np.float32(p["scoreThreshold"])) if isinstance(p.get("scoreThreshold"), float) else float(np.float32(0.7))
        self.nms_iou_threshold = float(np.float32(p["nmsIOUThreshold"])) if isinstance(p.get("nmsIOUThreshold"), float) else float(np.float32(0.3))
        c = self.ctx.cfg
        if c.max_detections != self.max_detections or c.detection_min_score != self.low_confidence_score_threshold \
                or c.detection_nms_iou != self.nms_iou_threshold:
            raise MaskRCNNError(_cabi.EINVAL, "DetectionLayer parameters differ from the context configuration")

    def output_shapes(self, for_input_shapes):
        rois = for_input_shapes[0]
        return [[self.max_detections, rois[1], 6, 1, 1]]

    def evaluate(self, inputs, outputs, keep_roi=None, count=None):
        rois, cls = inputs
        batch, _ = _batch_of(rois, 2)
        r = rois.shape[-2]
        check(self.ctx.handle, lib().mrcnn_detection_eval(self.ctx.handle, batch, r, ptr(rois), ptr(cls),
                                                          ptr(outputs[0]), ptr(keep_roi), ptr(count)))


class TimeDistributedMaskLayer(_Layer):
    """TimeDistributedMaskLayer.swift:14-92."""

    def output_shapes(self, for_input_shapes):
        s = for_input_shapes[0]
        return [[1, s[1], s[0], int(s[3]) * 2, int(s[4]) * 2]]

    def evaluate(self, inputs, outputs):
        pooled, detections = inputs
        batch, _ = _batch_of(pooled, 4)
        d = pooled.shape[-4]
        check(self.ctx.handle, lib().mrcnn_mask_eval(self.ctx.handle, batch, d, ptr(pooled), ptr(detections),
                                                     ptr(outputs[0])))
