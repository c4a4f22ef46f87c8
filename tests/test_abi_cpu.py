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


def test_header_is_plain_c_and_example_links(tmp_path):
    """include/maskrcnn_cuda.h is what a Swift module map / cgo / JNI stub consumes: it must compile as strict C99, and the
    C example (examples/predict.c: the streaming calls + decode) must link against the library.  Running it here (no GPU,
    no products) must fail loudly through the status / mrcnn_last_error path, never crash."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not available")
    exe = str(tmp_path / "predict_example")
    libdir = os.path.join(root, "mask-rcnn-coreml_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "examples", "predict.c"), "-L" + libdir, "-lmaskrcnn_cuda", "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=libdir + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe, str(tmp_path), "1", "1"], env=env, capture_output=True, text=True, timeout=120)
    import torch
    if not torch.cuda.is_available():
        assert r.returncode == 1 and "mrcnn_create failed" in r.stderr and len(r.stderr.strip()) > len("mrcnn_create failed (-2):")


def test_swift_and_cpp_bindings_only_use_declared_symbols():
    """The Swift sources cannot be compiled here (no toolchain): at least every mrcnn_* identifier they (and the C++
    mirror, which IS compiled) use must be declared in include/maskrcnn_cuda.h, every struct field they set must exist,
    and the five @objc layer classes of the reference must all be bound."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "maskrcnn_cuda.h")).read()
    declared = set(re.findall(r"\b(mrcnn_[a-z0-9_]+)\s*\(", header)) | {"mrcnn_config", "mrcnn_ctx", "mrcnn_status"}
    fields = set(re.findall(r"\b([a-z_0-9]+)(?:\[\d\])?\s*[;,]", header[header.index("typedef struct mrcnn_config"):header.index("} mrcnn_config;")]))
    swift_dir = os.path.join(root, "swift", "Sources", "MaskRCNNCuda")
    sources = {f: open(os.path.join(swift_dir, f)).read() for f in os.listdir(swift_dir) if f.endswith(".swift")}
    sources["maskrcnn.hpp"] = open(os.path.join(root, "include", "maskrcnn.hpp")).read()
    for name, text in sources.items():
        used = set(re.findall(r"\b(mrcnn_[a-z0-9_]+)\b", text)) - {"mrcnn_status"}
        assert used <= declared, f"{name} uses undeclared symbols: {sorted(used - declared)}"
        if name.endswith(".swift"):
            for f in re.findall(r"\bcfg\.([a-z_0-9]+)", text):
                assert f in fields, f"{name} sets mrcnn_config.{f}, which the header does not declare"
    objc = set(re.findall(r"@objc\((\w+)\)", "".join(t for n, t in sources.items() if n.endswith(".swift"))))
    assert objc == {"ProposalLayer", "PyramidROIAlignLayer", "TimeDistributedClassifierLayer", "DetectionLayer", "TimeDistributedMaskLayer"}


def test_anchor_generation_matches_the_python_generator():
    """mrcnn_generate_anchors (host arithmetic, the reference's "generate the anchors on demand" TODO) is bit-identical to
    synth.generate_anchors, i.e. to the anchors.bin content the tests and the bench use."""
    import ctypes as C
    import numpy as np
    import maskrcnn_b200 as m
    l = m.lib()
    for h, w, n in ((1024, 1024, 261888), (512, 512, 65472), (256, 256, 16368), (128, 128, 4092), (640, 384, None), (100, 70, None)):
        want = m.synth.generate_anchors(h, w)
        cnt = l.mrcnn_anchor_count(h, w)
        assert cnt == want.shape[0] and (n is None or cnt == n)
        got = np.full((cnt + 1, 4), 7.0, np.float32)
        assert l.mrcnn_generate_anchors(h, w, got.ctypes.data_as(C.c_void_p), cnt) == 0
        np.testing.assert_array_equal(got[:cnt], want)
        assert (got[cnt] == 7.0).all()                                   # nothing written past N rows
        assert l.mrcnn_generate_anchors(h, w, got.ctypes.data_as(C.c_void_p), cnt - 1) == m._cabi.EINVAL
    assert l.mrcnn_anchor_count(1, 1024) == m._cabi.EINVAL and l.mrcnn_generate_anchors(1024, 1024, None, 1 << 20) == m._cabi.EINVAL


def test_anchor_rule_is_the_upstream_meshgrid_construction():
    """The anchors (file content of anchors.bin, Conversion/task.py:176) against the upstream Matterport construction written
    the way that package writes it (meshgrids of scales x ratios and of shifts; norm_boxes), SURVEY.md Appendix B."""
    import numpy as np
    import maskrcnn_b200 as m

    def upstream(scales, ratios, shape, feature_stride, anchor_stride=1):
        scales, ratios = np.meshgrid(np.array(scales), np.array(ratios))
        scales, ratios = scales.flatten(), ratios.flatten()
        heights, widths = scales / np.sqrt(ratios), scales * np.sqrt(ratios)
        shifts_y = np.arange(0, shape[0], anchor_stride) * feature_stride
        shifts_x = np.arange(0, shape[1], anchor_stride) * feature_stride
        shifts_x, shifts_y = np.meshgrid(shifts_x, shifts_y)
        box_widths, box_centers_x = np.meshgrid(widths, shifts_x)
        box_heights, box_centers_y = np.meshgrid(heights, shifts_y)
        box_centers = np.stack([box_centers_y, box_centers_x], axis=2).reshape([-1, 2])
        box_sizes = np.stack([box_heights, box_widths], axis=2).reshape([-1, 2])
        return np.concatenate([box_centers - 0.5 * box_sizes, box_centers + 0.5 * box_sizes], axis=1)

    for h, w in ((1024, 1024), (512, 512), (640, 384)):
        boxes = np.concatenate([upstream(s, (0.5, 1, 2), (int(np.ceil(h / st)), int(np.ceil(w / st))), st)
                                for s, st in zip((32, 64, 128, 256, 512), (4, 8, 16, 32, 64))], axis=0)
        norm = ((boxes - np.array([0, 0, 1, 1])) / np.array([h - 1, w - 1, h - 1, w - 1])).astype(np.float32)     # norm_boxes
        np.testing.assert_array_equal(m.synth.generate_anchors(h, w), norm)
