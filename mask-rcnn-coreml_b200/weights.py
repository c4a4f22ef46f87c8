"""Weights of the model split (MaskRCNN / Classifier / Mask) for libmaskrcnn_cuda.so.

The reference ships its dense graphs as Core ML artefacts converted from Keras
(`Conversion/task.py:69-116`, fp16 weights `:90,102,114`); neither the artefacts
nor the Keras package are in the reference tree.  This
module defines

  * the layer table of the (Matterport) architecture the artefacts hold,
  * "reference layout" parameters: Keras-style HWIO conv kernels + bias + BatchNorm
    statistics, Dense kernels [in, out], Conv2DTranspose kernels [kh, kw, out, in],
  * fold(): BatchNorm folded into weight + bias, rounded to fp16 like the converter
    does, re-laid out for the tcgen05 implicit-GEMM kernel ([cout][kh][kw][cin]),
  * pack_blob()/unpack_blob(): the "MRCNNW1" file format that mrcnn_set_weights /
    the *_model_path configuration fields load,
  * synthetic(): seeded random parameters (no trained weights are reachable offline).

Blob format (little endian): header {magic "MRCNNW1\\0", u32 version=1, u32 which,
u32 n_tensors, u32 0}; n_tensors entries {char name[64], u32 dtype (0 f16, 1 f32),
u32 ndim, u64 dims[4], u64 offset, u64 nbytes}; payload, every tensor 256-byte aligned.
"""
import struct

import numpy as np

BN_EPS = 1e-3          # Keras BatchNormalization default epsilon
MAIN, CLASSIFIER, MASK = 0, 1, 2
_HEADER = struct.Struct("<8sIIII")
_ENTRY = struct.Struct("<64sII4QQQ")


def resnet_blocks(architecture):
    return {101: (3, 4, 23, 3), 50: (3, 4, 6, 3)}[architecture]


def layer_table(architecture=101, num_classes=81, pool_classifier=7):
    """[(name, which, kind, kh, kw, cin, cout, has_bn, relu)] in execution order."""
    t = [("conv1", MAIN, "conv", 7, 7, 3, 64, True, True)]
    cin = 64
    for s, nb in enumerate(resnet_blocks(architecture)):
        f = 64 << s
        for i in range(nb):
            p = f"res{s + 2}.{i}"
            t.append((p + ".2a", MAIN, "conv", 1, 1, cin, f, True, True))
            t.append((p + ".2b", MAIN, "conv", 3, 3, f, f, True, True))
            if i == 0:
                t.append((p + ".1", MAIN, "conv", 1, 1, cin, 4 * f, True, False))
            t.append((p + ".2c", MAIN, "conv", 1, 1, f, 4 * f, True, True))
            cin = 4 * f
    for l, c in zip((5, 4, 3, 2), (2048, 1024, 512, 256)):
        t.append((f"fpn.c{l}p{l}", MAIN, "conv", 1, 1, c, 256, False, False))
    for l in (2, 3, 4, 5):
        t.append((f"fpn.p{l}", MAIN, "conv", 3, 3, 256, 256, False, False))
    t.append(("rpn.shared", MAIN, "conv", 3, 3, 256, 512, False, True))
    t.append(("rpn.head", MAIN, "conv", 1, 1, 512, 18, False, False))       # 6 class logits (anchor, bg/fg) + 12 deltas
    P = pool_classifier
    t.append(("cls.conv1", CLASSIFIER, "conv", P, P, 256, 1024, True, True))
    t.append(("cls.conv2", CLASSIFIER, "conv", 1, 1, 1024, 1024, True, True))
    t.append(("cls.fc", CLASSIFIER, "conv", 1, 1, 1024, num_classes * 5, False, False))  # logits | (ncls,4) deltas
    for i in range(1, 5):
        t.append((f"mask.conv{i}", MASK, "conv", 3, 3, 256, 256, True, True))
    t.append(("mask.deconv", MASK, "deconv", 2, 2, 256, 256, False, True))
    t.append(("mask.final", MASK, "conv", 1, 1, 256, num_classes, False, False))
    return t


def synthetic(architecture=101, num_classes=81, pool_classifier=7, seed=7):
    """Seeded reference-layout parameters: {name: {"kernel", "bias", ["bn": (gamma, beta, mean, var)]}}.

    He-normal kernels; the last BatchNorm of every residual branch gets gamma ~0.3 so
    that activations stay inside fp16 range through 33 blocks; the classifier's logit rows get a larger
    gain so that soft-max scores spread over (0, 1) and the score filter / NMS paths see work."""
    params = {}
    for li, (name, which, kind, kh, kw, cin, cout, has_bn, relu) in enumerate(layer_table(architecture, num_classes, pool_classifier)):
        rng = np.random.default_rng([seed, li])
        fan_in = kh * kw * cin
        std = np.sqrt((2.0 if relu else 1.0) / fan_in)
        if kind == "deconv":
            kernel = rng.standard_normal((kh, kw, cout, cin), dtype=np.float32) * np.sqrt(2.0 / cin)
        else:
            kernel = rng.standard_normal((kh, kw, cin, cout), dtype=np.float32) * std
        if name == "conv1":
            kernel /= 74.0            # pixels minus mean have std ~74: bring activations to O(1)
        if name == "cls.fc":
            kernel[..., :num_classes] *= 2.0
            kernel[..., num_classes:] *= 0.5
        if name == "rpn.head":
            kernel[..., :6] *= 0.6
            kernel[..., 6:] *= 0.5
        p = {"kernel": kernel.astype(np.float32), "bias": (0.01 * rng.standard_normal(cout)).astype(np.float32)}
        if has_bn:
            g0 = 0.3 if name.endswith(".2c") else 1.0
            p["bn"] = ((g0 * (1.0 + 0.1 * rng.standard_normal(cout))).astype(np.float32),
                       (0.05 * rng.standard_normal(cout)).astype(np.float32),
                       (0.05 * rng.standard_normal(cout)).astype(np.float32),
                       (1.0 + 0.1 * np.abs(rng.standard_normal(cout))).astype(np.float32))
        params[name] = p
    return params


def fold(params, architecture=101, num_classes=81, pool_classifier=7):
    """Reference-layout parameters -> {name: (w [cout,kh,kw,cin] float16, b [cout] float32)} with BN folded in.

    Rounding to fp16 happens after folding (the converter quantises the folded Core ML weights,
    Conversion/task.py:90)."""
    out = {}
    for name, which, kind, kh, kw, cin, cout, has_bn, relu in layer_table(architecture, num_classes, pool_classifier):
        p = params[name]
        k = p["kernel"].astype(np.float64)
        b = p["bias"].astype(np.float64)
        if kind == "deconv":
            w = k.reshape(kh * kw * cout, 1, 1, cin)        # row (dy*2+dx)*cout + co : out[2y+dy, 2x+dx, co]
            b = np.tile(b, kh * kw)
        else:
            w = np.transpose(k, (3, 0, 1, 2))               # HWIO -> OHWI
        if "bn" in p:                 # (imported artefacts may come with the batch norm already fused: no "bn" then)
            gamma, beta, mean, var = [np.asarray(x).astype(np.float64) for x in p["bn"]]
            scale = gamma / np.sqrt(var + float(p.get("bn_eps", BN_EPS)))
            w = w * scale[:, None, None, None]
            b = (b - mean) * scale + beta
        out[name] = (np.ascontiguousarray(w).astype(np.float16), b.astype(np.float32))
    return out


def pack_conv1_s2d(w):
    """[64,7,7,3] folded stem kernel -> [64,4,1,64]: the 7x7/2 convolution over the zero-padded image as a
    4-tap (dy) GEMM over the 2x2 space-to-depth image, K = dy*64 + dx*16 + (p*2+q)*3 + c with
    (ky, kx) = (2dy+p, 2dx+q); positions with ky or kx == 7 and channels 12..15 are zero."""
    o = np.zeros((w.shape[0], 4, 1, 64), dtype=np.float16)
    for dy in range(4):
        for p in range(2):
            ky = 2 * dy + p
            if ky >= 7:
                continue
            for dx in range(4):
                for q in range(2):
                    kx = 2 * dx + q
                    if kx >= 7:
                        continue
                    base = dx * 16 + (p * 2 + q) * 3
                    o[:, dy, 0, base:base + 3] = w[:, ky, kx, :]
    return o


def device_tensors(folded, which, architecture=101, num_classes=81, pool_classifier=7):
    """The tensors of one blob, in table order: [(name, array)]."""
    ts = []
    for name, wh, *_ in layer_table(architecture, num_classes, pool_classifier):
        if wh != which:
            continue
        w, b = folded[name]
        if name == "conv1":
            w = pack_conv1_s2d(w)
        ts.append((name + ".w", w))
        ts.append((name + ".b", b))
    return ts


def pack_blob(which, tensors):
    n = len(tensors)
    off = _HEADER.size + n * _ENTRY.size
    entries, chunks = [], []
    for name, arr in tensors:
        arr = np.ascontiguousarray(arr)
        dtype = {np.dtype(np.float16): 0, np.dtype(np.float32): 1}[arr.dtype]
        pad = (-off) % 256
        chunks.append(b"\0" * pad)
        off += pad
        dims = list(arr.shape) + [0] * (4 - arr.ndim)
        entries.append(_ENTRY.pack(name.encode(), dtype, arr.ndim, *dims, off, arr.nbytes))
        chunks.append(arr.tobytes())
        off += arr.nbytes
    return _HEADER.pack(b"MRCNNW1\0", 1, which, n, 0) + b"".join(entries) + b"".join(chunks)


def unpack_blob(blob):
    magic, version, which, n, _ = _HEADER.unpack_from(blob, 0)
    assert magic == b"MRCNNW1\0" and version == 1
    out = {}
    for i in range(n):
        name, dtype, ndim, d0, d1, d2, d3, off, nbytes = _ENTRY.unpack_from(blob, _HEADER.size + i * _ENTRY.size)
        dt = np.float16 if dtype == 0 else np.float32
        out[name.rstrip(b"\0").decode()] = np.frombuffer(blob, dtype=dt, count=nbytes // np.dtype(dt).itemsize,
                                                         offset=off).reshape([d0, d1, d2, d3][:ndim])
    return which, out


def synthetic_blobs(architecture=101, num_classes=81, pool_classifier=7, seed=7):
    """(folded, [main_blob, classifier_blob, mask_blob]) for seeded synthetic weights."""
    folded = fold(synthetic(architecture, num_classes, pool_classifier, seed), architecture, num_classes, pool_classifier)
    return folded, [pack_blob(w, device_tensors(folded, w, architecture, num_classes, pool_classifier)) for w in (MAIN, CLASSIFIER, MASK)]


def write_products(directory, architecture=101, image_h=1024, image_w=1024, num_classes=81, seed=7):
    """Writes the model split the reference keeps under .maskrcnn/models/<name>/products/ (README.md:107-116):
    anchors.bin + MaskRCNN / Classifier / Mask (here as .mrcnnw blobs).  Returns the four paths."""
    import os
    from . import synth
    os.makedirs(directory, exist_ok=True)
    _, blobs = synthetic_blobs(architecture, num_classes, 7, seed)
    paths = {}
    for name, blob in zip(("MaskRCNN", "Classifier", "Mask"), blobs):
        paths[name] = os.path.join(directory, name + ".mrcnnw")
        with open(paths[name], "wb") as f:
            f.write(blob)
    paths["anchors"] = os.path.join(directory, "anchors.bin")
    synth.generate_anchors(image_h, image_w).tofile(paths["anchors"])      # Conversion/task.py:176
    return paths
