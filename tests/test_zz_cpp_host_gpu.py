"""GPU parity of the C++ host mirror (include/maskrcnn.hpp): tests/cpp/host_mirror_gpu.cpp drives the reference-shaped
C++ classes on the inputs of the committed golden fixtures; every output must equal the golden arrays bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from test_cpp_host_cpu import ROOT, build_cpp

pytestmark = pytest.mark.gpu
G = os.path.join(ROOT, "tests", "golden")


def test_cpp_host_mirror_golden_gpu(tmp_path, orc):
    exe = str(tmp_path / "host_mirror_gpu")
    env = build_cpp(os.path.join(ROOT, "tests", "cpp", "host_mirror_gpu.cpp"), exe)
    d = tmp_path / "io"
    d.mkdir()
    p, ra, dt, md = (np.load(os.path.join(G, n + ".npz")) for n in ("proposal", "roialign", "detection", "mask_decode"))

    def put(name, a, dtype=np.float32):
        np.ascontiguousarray(a, dtype=dtype).tofile(str(d / name))

    put("anchors.f32", p["anchors"]); put("probs.f32", p["probs"]); put("tie_probs.f32", p["tie_probs"]); put("deltas.f32", p["deltas"])
    put("ra_rois.f32", ra["rois"])
    for l in range(4):
        put(f"maps{l}.f32", ra[f"maps{l}"])
    put("cls_probs.f32", dt["probs"]); put("cls_bbox.f32", dt["bbox"]); put("det_rois.f32", dt["rois"]); put("det_cls.f32", dt["cls"])
    put("masks.f32", md["masks"])
    r = subprocess.run([exe, str(d)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "kernels launched" in r.stdout

    def get(name, dtype, shape):
        return np.fromfile(str(d / name), dtype=dtype).reshape(shape)

    for tag in ("", "tie_"):
        assert get(tag + "count.out", np.int32, (1,))[0] == int(p[tag + "count"])
        np.testing.assert_array_equal(get(tag + "keep.out", np.int32, (100,)), p[tag + "keep"])
        np.testing.assert_array_equal(get(tag + "rois.out", np.float32, (100, 4)), p[tag + "rois"])
    np.testing.assert_array_equal(get("levels.out", np.int32, (64,)), ra["levels"])
    np.testing.assert_array_equal(get("pooled7.out", np.float32, (64, 8, 7, 7)), ra["pooled7"])
    np.testing.assert_array_equal(get("pooled14.out", np.float32, (64, 8, 14, 14)), ra["pooled14"])
    np.testing.assert_array_equal(get("cls_select.out", np.float32, (200, 6)), orc.classifier_select(dt["probs"], dt["bbox"]))
    assert get("det_count.out", np.int32, (1,))[0] == int(dt["count"])
    np.testing.assert_array_equal(get("det_keep.out", np.int32, (100,)), dt["keep"])
    np.testing.assert_array_equal(get("det.out", np.float32, (100, 6)), dt["det"])
    n = int(md["n"])
    meta = get("dec_meta.out", np.int32, (-1, 2))
    real = get("dec_real.out", np.float64, (-1, 5))
    assert meta.shape[0] == n and real.shape[0] == n
    np.testing.assert_array_equal(meta[:, 0], md["index"][:n])
    np.testing.assert_array_equal(meta[:, 1], md["classes"][:n])
    np.testing.assert_array_equal(real[:, :4], md["bbox"][:n])
    np.testing.assert_array_equal(real[:, 4], md["score"][:n])
    np.testing.assert_array_equal(get("dec_mask.out", np.uint8, (n, 784)), md["mask_u8"][:n])
