"""maskrcnn_b200 -- B200-native drop-in for the Sources/Mask-RCNN-CoreML path.

The directory is named ``mask-rcnn-coreml_b200`` (not importable by name); import
it through the repo-root shim ``import maskrcnn_b200``.
"""
from ._cabi import MaskRCNNError, LIB_PATH, lib  # noqa: F401
from .layers import (Context, DetectionLayer, ProposalLayer, PyramidROIAlignLayer,  # noqa: F401
                     TimeDistributedClassifierLayer, TimeDistributedMaskLayer, default_context)
from .model import Detection, MaskRCNN, MaskRCNNConfig  # noqa: F401
from . import coco, distributed, h5lite, keras_h5, mlmodel, results_pb, synth, weights  # noqa: F401
