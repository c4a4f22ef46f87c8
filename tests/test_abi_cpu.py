"""CPU tests of the boundary: the shared library loads, exports every symbol that
include/maskrcnn_cuda.h declares, and fails loudly without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "maskrcnn_cuda.h")).read()
    return sorted(set(re.findall(r"MRCNN_API[^;(]*?\b(mrcnn_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    names = _declared_symbols()
    assert len(names) >= 25
    l = C.CDLL(pkg.LIB_PATH)
    missing = [n for n in names if not hasattr(l, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    # and the ctypes table binds exactly the declared set
    assert sorted(pkg._cabi.SIGNATURES) == names


def test_config_defaults_match_reference(pkg):
    cfg = pkg._cabi.mrcnn_config()
    pkg.lib().mrcnn_config_default(C.byref(cfg))
    assert cfg.struct_size == C.sizeof(pkg._cabi.mrcnn_config)
    assert (cfg.image_h, cfg.image_w, cfg.architecture, cfg.num_classes) == (1024, 1024, 101, 81)
    assert [round(v, 6) for v in cfg.bbox_std] == [0.1, 0.1, 0.2, 0.2]           # ProposalLayer.swift:57
    assert (cfg.pre_nms_max_proposals, cfg.max_proposals) == (6000, 1000)        # :59,:61
    assert cfg.proposal_nms_iou == C.c_float(0.7).value                          # :63
    assert (cfg.pool_size_classifier, cfg.pool_size_mask) == (7, 14)
    assert cfg.max_detections == 100 and cfg.detection_min_score == C.c_float(0.7).value
    assert cfg.detection_nms_iou == C.c_float(0.3).value                         # DetectionLayer.swift:61
    assert [round(v, 4) for v in cfg.mean_rgb] == [123.7, 116.8, 103.9]          # Conversion/task.py:73-75
    assert pkg.lib().mrcnn_version().startswith(b"maskrcnn_cuda")


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.MaskRCNNError) as e:
        pkg.Context()
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    pkg_dir = os.path.join(ROOT, "mask-rcnn-coreml_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".swift")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in src and "from oracle" not in src and "import oracle" not in src, f
