"""h5lite (HDF5 subset reader / writer) and the Keras-weights importer.  No HDF5 library or file is
reachable offline: these tests check the reader against files built by the module's own, separately written encoder
(old-style groups with multi-level B-trees, continuation blocks, layout versions 1 and 3, chunked + shuffled + deflated
data) and against a hand-assembled file -- self-consistency with the format specification as understood here."""
import struct

import numpy as np
import pytest


@pytest.fixture(scope="module")
def m():
    import maskrcnn_b200
    return maskrcnn_b200


def _tree(rng):
    return {
        "conv1": {"conv1": {"kernel:0": rng.standard_normal((7, 7, 3, 8)).astype(np.float32), "bias:0": rng.standard_normal(8).astype(np.float32)}},
        "rpn_model": {"rpn_conv_shared": {"kernel:0": rng.standard_normal((3, 3, 4, 6)).astype(np.float32)}},
        "misc": {"f64": rng.standard_normal((5, 3)), "f16": rng.standard_normal(9).astype(np.float16),
                 "i32": rng.integers(-5, 5, (2, 3, 4)).astype(np.int32), "u8": rng.integers(0, 255, 17).astype(np.uint8),
                 "scalar": np.float32(2.5), "empty": np.zeros((0, 4), np.float32)},
    }


def _flat(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(_flat(v, prefix + k + "/"))
        else:
            out[prefix + k] = np.asarray(v)
    return out


@pytest.mark.parametrize("kw", [dict(), dict(layout_version=1), dict(split_headers=True),
                                dict(chunks=lambda a: tuple(max(1, (s + 1) // 2) for s in a.shape) if a.ndim and a.size else None),
                                dict(chunks=lambda a: tuple(max(1, (s + 2) // 3) for s in a.shape) if a.ndim and a.size else None, deflate=True, shuffle=True)])
def test_round_trip_every_storage_kind(m, kw):
    tree = _tree(np.random.default_rng(0))
    f = m.h5lite.read(m.h5lite.write(tree, **kw))
    got = f.datasets()
    want = _flat(tree)
    assert sorted(got) == sorted(want)
    for k, a in want.items():
        assert got[k].dtype == a.dtype and got[k].shape == a.shape, k
        np.testing.assert_array_equal(got[k], a)
    np.testing.assert_array_equal(f["conv1/conv1/bias:0"], tree["conv1"]["conv1"]["bias:0"])
    assert "/rpn_model/rpn_conv_shared" in f.groups()
    with pytest.raises(KeyError):
        f["conv1/nope"]
    with pytest.raises(KeyError):
        f["conv1/conv1"]


def test_many_links_use_a_multi_level_btree(m):
    """~400 layer groups (ResNet101 Mask R-CNN has 394) with leaf K = 4, internal K = 16: 50 symbol-table nodes under a
    2-level B-tree; a large chunked dataset needs a 2-level chunk B-tree as well."""
    rng = np.random.default_rng(1)
    tree = {f"layer_{i:03d}": {"w": rng.standard_normal(3).astype(np.float32)} for i in range(400)}
    big = rng.standard_normal((40, 40)).astype(np.float32)
    tree["big"] = big
    data = m.h5lite.write(tree, chunks=lambda a: (4, 4) if a.ndim == 2 else None)
    got = m.h5lite.read(data).datasets()
    assert len(got) == 401
    for i in (0, 7, 8, 255, 256, 399):
        np.testing.assert_array_equal(got[f"layer_{i:03d}/w"], tree[f"layer_{i:03d}"]["w"])
    np.testing.assert_array_equal(got["big"], big)


def test_hand_assembled_file(m):
    """A file put together byte by byte from the format specification, independent of h5lite.write: superblock v0, root
    group (symbol-table message only: cache type 0), one contiguous 2x3 big-endian float64 dataset 'd'."""
    buf = bytearray(96)
    def at(pos, b):
        if len(buf) < pos + len(b):
            buf.extend(bytes(pos + len(b) - len(buf)))
        buf[pos:pos + len(b)] = b
    data = np.arange(6, dtype=">f8").reshape(2, 3)
    at(0x400, data.tobytes())
    # dataset object header @0x200
    space = struct.pack("<BBBB4xQQ", 1, 2, 0, 0, 2, 3)
    dtype = struct.pack("<BBBBI", 0x11, 0x21, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    layout = struct.pack("<BBQQ", 3, 1, 0x400, 48)
    def msg(t, body):
        body += bytes(-len(body) % 8)
        return struct.pack("<HHB3x", t, len(body), 0) + body
    msgs = msg(1, space) + msg(3, dtype) + msg(0, b"") + msg(8, layout)          # incl. a NIL message
    at(0x200, struct.pack("<BxHII4x", 1, 4, 1, len(msgs)) + msgs)
    # local heap @0x300 (data segment @0x340): "\0" then "d\0"
    at(0x340, b"\0" * 8 + b"d\0" + bytes(6))
    at(0x300, b"HEAP" + struct.pack("<B3xQQQ", 0, 16, 0xFFFFFFFFFFFFFFFF, 0x340))
    # symbol table node @0x500 with one entry, B-tree @0x600
    at(0x500, b"SNOD" + struct.pack("<BxH", 1, 1) + struct.pack("<QQII16x", 8, 0x200, 0, 0))
    at(0x600, b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF) + struct.pack("<QQQ", 0, 0x500, 8))
    # root group object header @0x100
    rmsg = msg(0x11, struct.pack("<QQ", 0x600, 0x300))
    at(0x100, struct.pack("<BxHII4x", 1, 1, 1, len(rmsg)) + rmsg)
    sb = m.h5lite.SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, 0xFFFFFFFFFFFFFFFF, len(buf), 0xFFFFFFFFFFFFFFFF) + struct.pack("<QQII16x", 0, 0x100, 0, 0)
    at(0, sb)
    f = m.h5lite.read(bytes(buf))
    got = f["d"]
    assert got.dtype == np.float64 and got.shape == (2, 3)
    np.testing.assert_array_equal(got, np.arange(6, dtype=np.float64).reshape(2, 3))
    # the same file behind a 512-byte user block, addresses relative to the base address
    shifted = bytearray(512) + buf
    shifted[512 + 24:512 + 32] = struct.pack("<Q", 512)
    np.testing.assert_array_equal(m.h5lite.read(bytes(shifted))["d"], got)


def test_unsupported_and_damaged_files_fail_loudly(m):
    H = m.h5lite
    with pytest.raises(H.H5Error, match="not an HDF5 file"):
        H.read(b"PK\x03\x04" + bytes(600))
    good = H.write({"g": {"w": np.arange(12, dtype=np.float32).reshape(3, 4), "c": np.arange(64, dtype=np.float32).reshape(8, 8)}},
                   chunks=lambda a: (4, 4) if a.size == 64 else None, deflate=True, shuffle=True)
    v2 = bytearray(good)
    v2[8] = 2
    with pytest.raises(H.H5Error, match="superblock version 2"):
        H.read(bytes(v2))
    rng = np.random.default_rng(5)
    survived = 0
    for _ in range(2000):                                # truncations and byte flips: H5Error / KeyError or a clean read, never a crash
        b = bytearray(good)
        if rng.random() < 0.3:
            b = b[:int(rng.integers(8, len(b)))]
        else:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(8, len(b)))] = int(rng.integers(0, 256))
        try:
            H.read(bytes(b)).datasets()
            survived += 1
        except (H.H5Error, KeyError):
            pass
    assert survived < 2000


@pytest.mark.parametrize("arch", [50])
def test_keras_weights_import_gives_the_same_blobs(m, arch):
    """synthetic reference-layout parameters -> Keras save_weights tree -> HDF5 bytes -> importer: the three weight blobs
    must be byte-identical to those packed from the parameters directly (BN folding, Dense -> 1x1, rpn / classifier
    head concatenation, deconvolution layout all on the path)."""
    params = m.weights.synthetic(arch)
    tree = m.keras_h5.keras_tree_from_params(params, arch)
    assert "rpn_conv_shared" in tree["rpn_model"] and tree["mrcnn_bbox_fc"]["mrcnn_bbox_fc"]["kernel:0"].shape == (1024, 324)
    data = m.h5lite.write(tree)
    blobs = m.keras_h5.import_products(data, arch)
    _, want = m.weights.synthetic_blobs(arch)
    assert [len(b) for b in blobs] == [len(b) for b in want]
    for got, ref in zip(blobs, want):
        assert got == ref
    broken = {k: v for k, v in tree.items() if k != "bn2a_branch2b"}
    with pytest.raises(KeyError, match="bn2a_branch2b"):
        m.keras_h5.params_from_keras_h5(m.h5lite.write(broken), arch)


def test_import_cli_writes_products(m, tmp_path):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "weights.h5"
    src.write_bytes(m.h5lite.write(m.keras_h5.keras_tree_from_params(m.weights.synthetic(50), 50)))
    out = tmp_path / "products"
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "import_keras_h5.py"), "--weights", str(src), "--architecture", "50",
                        "--image-size", "512", "--out", str(out), "--anchors"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    _, want = m.weights.synthetic_blobs(50)
    for name, ref in zip(("MaskRCNN", "Classifier", "Mask"), want):
        assert (out / (name + ".mrcnnw")).read_bytes() == ref
    assert (out / "anchors.bin").stat().st_size == 65472 * 16          # config S: 512x512 -> 65 472 anchors
