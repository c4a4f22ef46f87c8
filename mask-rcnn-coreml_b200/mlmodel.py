"""import the reference's real artefacts -- the three Core ML models of the model split
(`MaskRCNN.mlmodel`, `Classifier.mlmodel`, `Mask.mlmodel`, README.md:107-116, written by Conversion/task.py:69-116) --
into the MRCNNW1 blobs libmaskrcnn_cuda.so loads, without coremltools / protobuf-generated classes (neither is in the
image): a small protobuf wire-format reader plus the handful of Core ML `Model.proto` / `NeuralNetwork.proto` fields
that carry weights.

PARITY UNPINNED.  No `.mlmodel` of the reference is reachable offline, so nothing here has been run against a real
artefact.  What is pinned: the wire encoding (cross-checked against google.protobuf in tests/test_mlmodel_cpu.py) and
the round trip write_mlmodel -> read -> fold == weights.fold.  The field numbers below are those of the public Core ML
specification (coremltools `mlmodel/format/*.proto`, specification version 2-3 era = the converter of
Conversion/requirements.txt:3); the layer names are the Keras layer names of the credited Matterport graph, which the
Keras converter copies into the Core ML layer names.

    Model            specificationVersion=1  description=2  neuralNetwork=500 (also 303 regressor / 403 classifier)
    NeuralNetwork    layers=1  preprocessing=2
    NeuralNetworkLayer   name=1 input=2 output=3  convolution=100 innerProduct=140 batchnorm=160 custom=500
    ConvolutionLayerParams  outputChannels=1 kernelChannels=2 nGroups=10 kernelSize=20 stride=30 dilationFactor=40
                            valid=50 same=51 isDeconvolution=60 hasBias=70 weights=90 bias=91 outputShape=100
    InnerProductLayerParams inputChannels=1 outputChannels=2 hasBias=10 weights=20 bias=21
    BatchnormLayerParams    channels=1 computeMeanVar=5 instanceNormalization=6 epsilon=10 gamma=15 beta=16 mean=17 variance=18
    WeightParams            floatValue=1 (packed f32) float16Value=2 (bytes) rawValue=30
    CustomLayerParams       className=10 weights=20 parameters=30 (map<string, CustomLayerParamValue>) description=40
    CustomLayerParamValue   doubleValue=10 stringValue=20 intValue=30 longValue=40 boolValue=50
    NeuralNetworkPreprocessing  featureName=1 scaler=10 {channelScale=10 blueBias=20 greenBias=21 redBias=22 grayBias=30}

Kernel layouts: Core ML convolution weights are [Cout, Cin/groups, kH, kW] (deconvolution: [Cin, Cout/groups, kH, kW]),
inner-product weights [out, in]; weights.py's reference layout is Keras' (HWIO, Conv2DTranspose [kh, kw, out, in]).
"""
import struct

import numpy as np

from . import weights as W

# ---------------------------------------------------------------------------------------------------------------
# protobuf wire format
# ---------------------------------------------------------------------------------------------------------------
VARINT, I64, LEN, I32 = 0, 1, 2, 5


def _read_varint(buf, pos):
    result = shift = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def parse_message(buf):
    """bytes -> [(field_number, wire_type, value)]; value = int (VARINT / I64 / I32 raw bits) or memoryview (LEN)."""
    buf = memoryview(buf)
    out, pos, n = [], 0, len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == VARINT:
            v, pos = _read_varint(buf, pos)
        elif wt == I64:
            if pos + 8 > n:
                raise ValueError("truncated 64-bit field")
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == LEN:
            ln, pos = _read_varint(buf, pos)
            if pos + ln > n:
                raise ValueError("truncated length-delimited field")
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == I32:
            if pos + 4 > n:
                raise ValueError("truncated 32-bit field")
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        out.append((field, wt, v))
    return out


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field, wt):
    return _varint((field << 3) | wt)


def enc_varint(field, v):
    return _key(field, VARINT) + _varint(int(v))


def enc_bytes(field, b):
    b = bytes(b)
    return _key(field, LEN) + _varint(len(b)) + b


def enc_string(field, s):
    return enc_bytes(field, s.encode())


def enc_float(field, v):
    return _key(field, I32) + struct.pack("<f", v)


def enc_double(field, v):
    return _key(field, I64) + struct.pack("<d", v)


def _first(fields, number, default=None):
    for f, _, v in fields:
        if f == number:
            return v
    return default


def _all(fields, number):
    return [v for f, _, v in fields if f == number]


def _uint(fields, number, default=0):
    """first varint occurrence of a scalar field (other wire types under that number are ignored)."""
    for f, wt, v in fields:
        if f == number and wt == VARINT:
            return v
    return default


def _float(fields, number, default=None):
    """first 32-bit occurrence of a float field."""
    for f, wt, v in fields:
        if f == number and wt == I32:
            return _f32(v)
    return default


def _text(fields, number):
    """first length-delimited occurrence of a string field, decoded ('' when absent or of another wire type)."""
    for f, wt, v in fields:
        if f == number and wt == LEN:
            return bytes(v).decode()
    return ""


def _sub(fields, number):
    """first length-delimited occurrence of a sub-message field, or None (a scalar with that number is not a message)."""
    for f, wt, v in fields:
        if f == number and wt == LEN:
            return v
    return None


def _repeated_u64(fields, number):
    """repeated uint64: packed (LEN) or unpacked (VARINT) encodings."""
    out = []
    for f, wt, v in fields:
        if f != number:
            continue
        if wt == VARINT:
            out.append(v)
        elif wt == LEN:
            pos = 0
            while pos < len(v):
                x, pos = _read_varint(v, pos)
                out.append(x)
    return out


def _f32(bits):
    return struct.unpack("<f", struct.pack("<I", bits))[0]


# ---------------------------------------------------------------------------------------------------------------
# Core ML messages -> python
# ---------------------------------------------------------------------------------------------------------------
def _weights(msg):
    """WeightParams -> float32 ndarray (1-D), or None."""
    if msg is None:
        return None
    f = parse_message(msg)
    h = _sub(f, 2)
    if h is not None and len(h):
        if len(h) % 2:
            raise ValueError("float16Value of odd length")
        return np.frombuffer(h, dtype="<f2").astype(np.float32)
    vals = []
    for num, wt, v in f:
        if num != 1:
            continue
        if wt == LEN:
            if len(v) % 4:
                raise ValueError("packed floatValue of a length that is not a multiple of 4")
            vals.append(np.frombuffer(v, dtype="<f4"))
        elif wt == I32:
            vals.append(np.array([_f32(v)], np.float32))
    if vals:
        return np.concatenate(vals).astype(np.float32)
    raw = _sub(f, 30)
    if raw is not None and len(raw):
        raise ValueError("quantised (rawValue) weights are not supported")
    return None


def _layer(msg):
    f = parse_message(msg)
    L = {"name": _text(f, 1), "inputs": [bytes(v).decode() for n, wt, v in f if n == 2 and wt == LEN],
         "outputs": [bytes(v).decode() for n, wt, v in f if n == 3 and wt == LEN], "type": "other"}
    conv, ip, bn, custom = _sub(f, 100), _sub(f, 140), _sub(f, 160), _sub(f, 500)
    if conv is not None:
        c = parse_message(conv)
        ks = _repeated_u64(c, 20) or [3, 3]
        st = _repeated_u64(c, 30) or [1, 1]
        cout, kc, groups = _uint(c, 1), _uint(c, 2), _uint(c, 10, 1) or 1
        deconv = bool(_uint(c, 60))
        w = _weights(_sub(c, 90))
        if len(ks) < 2 or len(st) < 2:
            raise ValueError("convolution without a 2-D kernel size / stride")
        shape = (kc, cout // groups, ks[0], ks[1]) if deconv else (cout, kc, ks[0], ks[1])
        L.update(type="deconvolution" if deconv else "convolution", cout=cout, kernel_channels=kc, groups=groups,
                 kernel_size=tuple(ks), stride=tuple(st), padding="same" if _sub(c, 51) is not None else "valid",
                 weights=None if w is None else w.reshape(shape),
                 bias=_weights(_sub(c, 91)) if _uint(c, 70) else None)
    elif ip is not None:
        c = parse_message(ip)
        cin, cout = _uint(c, 1), _uint(c, 2)
        w = _weights(_sub(c, 20))
        L.update(type="innerProduct", cin=cin, cout=cout, weights=None if w is None else w.reshape(cout, cin),
                 bias=_weights(_sub(c, 21)) if _uint(c, 10) else None)
    elif bn is not None:
        c = parse_message(bn)
        L.update(type="batchnorm", channels=_uint(c, 1), epsilon=_float(c, 10, 1e-5),
                 gamma=_weights(_sub(c, 15)), beta=_weights(_sub(c, 16)), mean=_weights(_sub(c, 17)),
                 variance=_weights(_sub(c, 18)))
    elif custom is not None:
        c = parse_message(custom)
        params = {}
        for n30, wt30, entry in c:                # map<string, CustomLayerParamValue> = repeated {key=1, value=2}
            if n30 != 30 or wt30 != LEN:
                continue
            e = parse_message(entry)
            k = _text(e, 1)
            v = parse_message(_sub(e, 2) or b"")
            for num, wt, x in v:
                if num == 10 and wt == I64:
                    params[k] = struct.unpack("<d", struct.pack("<Q", x))[0]
                elif num == 20 and wt == LEN:
                    params[k] = bytes(x).decode()
                elif num in (30, 40) and wt == VARINT:
                    params[k] = x - (1 << 64) if x >> 63 else x
                elif num == 50 and wt == VARINT:
                    params[k] = bool(x)
        L.update(type="custom", class_name=_text(c, 10), parameters=params)
    return L


def read_mlmodel(data):
    """bytes (or a path) of a .mlmodel -> {"specification_version", "layers": [...], "preprocessing": {...}}."""
    if isinstance(data, str):
        with open(data, "rb") as fh:
            data = fh.read()
    top = parse_message(data)
    nn = None
    for num in (500, 403, 303):
        nn = _sub(top, num)
        if nn is not None:
            break
    if nn is None:
        raise ValueError("not a neural-network Core ML model (no field 500 / 403 / 303)")
    f = parse_message(nn)
    pre = {}
    for n2, wt2, p in f:
        if n2 != 2 or wt2 != LEN:
            continue
        sc = _sub(parse_message(p), 10)
        if sc is not None:
            s = parse_message(sc)
            g = lambda n, d: _float(s, n, d)
            pre = {"channelScale": g(10, 1.0), "blueBias": g(20, 0.0), "greenBias": g(21, 0.0), "redBias": g(22, 0.0)}
    return {"specification_version": _uint(top, 1), "layers": [_layer(m) for n1, wt1, m in f if n1 == 1 and wt1 == LEN],
            "preprocessing": pre}


# ---------------------------------------------------------------------------------------------------------------
# writer (tests, and export of this repo's weights for Core ML tooling)
# ---------------------------------------------------------------------------------------------------------------
def _enc_weights(arr, half=True):
    arr = np.ascontiguousarray(arr, dtype=np.float32).ravel()
    if half:
        return enc_bytes(2, arr.astype("<f2").tobytes())
    return enc_bytes(1, arr.astype("<f4").tobytes())          # packed repeated float


def _enc_layer(L, half):
    body = enc_string(1, L["name"])
    for x in L.get("inputs", []):
        body += enc_string(2, x)
    for x in L.get("outputs", []):
        body += enc_string(3, x)
    t = L["type"]
    if t in ("convolution", "deconvolution"):
        w = L["weights"]
        c = enc_varint(1, L["cout"]) + enc_varint(2, L["kernel_channels"]) + enc_varint(10, 1)
        c += enc_bytes(20, b"".join(_varint(k) for k in L["kernel_size"])) + enc_bytes(30, b"".join(_varint(k) for k in L["stride"]))
        c += enc_bytes(51 if L.get("padding") == "same" else 50, b"")
        if t == "deconvolution":
            c += enc_varint(60, 1)
        if L.get("bias") is not None:
            c += enc_varint(70, 1)
        c += enc_bytes(90, _enc_weights(w, half))
        if L.get("bias") is not None:
            c += enc_bytes(91, _enc_weights(L["bias"], half))
        body += enc_bytes(100, c)
    elif t == "innerProduct":
        c = enc_varint(1, L["cin"]) + enc_varint(2, L["cout"])
        if L.get("bias") is not None:
            c += enc_varint(10, 1)
        c += enc_bytes(20, _enc_weights(L["weights"], half))
        if L.get("bias") is not None:
            c += enc_bytes(21, _enc_weights(L["bias"], half))
        body += enc_bytes(140, c)
    elif t == "batchnorm":
        c = enc_varint(1, L["channels"]) + enc_float(10, L["epsilon"])
        for num, k in ((15, "gamma"), (16, "beta"), (17, "mean"), (18, "variance")):
            c += enc_bytes(num, _enc_weights(L[k], half))
        body += enc_bytes(160, c)
    elif t == "custom":
        c = enc_string(10, L["class_name"])
        for k, v in L.get("parameters", {}).items():
            if isinstance(v, bool):
                val = enc_varint(50, int(v))
            elif isinstance(v, int):
                val = enc_varint(30, v & ((1 << 64) - 1))
            elif isinstance(v, float):
                val = enc_double(10, v)
            else:
                val = enc_string(20, str(v))
            c += enc_bytes(30, enc_string(1, k) + enc_bytes(2, val))
        body += enc_bytes(500, c)
    return body


def write_mlmodel(layers, preprocessing=None, half=True, specification_version=2):
    """[layer dict] (the shape read_mlmodel returns) -> .mlmodel bytes (weights as float16Value when `half`)."""
    nn = b"".join(enc_bytes(1, _enc_layer(L, half)) for L in layers)
    if preprocessing:
        sc = enc_float(10, preprocessing.get("channelScale", 1.0)) + enc_float(20, preprocessing.get("blueBias", 0.0))
        sc += enc_float(21, preprocessing.get("greenBias", 0.0)) + enc_float(22, preprocessing.get("redBias", 0.0))
        nn += enc_bytes(2, enc_string(1, "image") + enc_bytes(10, sc))
    return enc_varint(1, specification_version) + enc_bytes(500, nn)


# ---------------------------------------------------------------------------------------------------------------
# Matterport / Keras layer names <-> this repo's layer table
# ---------------------------------------------------------------------------------------------------------------
def keras_names(architecture=101):
    """{internal name: (conv layer name(s), batch-norm layer name or None)} for weights.layer_table()."""
    m = {"conv1": (["conv1"], "bn_conv1")}
    for s, nb in enumerate(W.resnet_blocks(architecture)):
        for i in range(nb):
            blk = "a" if i == 0 else chr(ord("b") + i - 1)          # conv_block 'a', identity blocks 'b', 'c', ...
            for br in ("2a", "2b", "2c"):
                m[f"res{s + 2}.{i}.{br}"] = ([f"res{s + 2}{blk}_branch{br}"], f"bn{s + 2}{blk}_branch{br}")
            if i == 0:
                m[f"res{s + 2}.0.1"] = ([f"res{s + 2}a_branch1"], f"bn{s + 2}a_branch1")
    for l in (2, 3, 4, 5):
        m[f"fpn.c{l}p{l}"] = ([f"fpn_c{l}p{l}"], None)
        m[f"fpn.p{l}"] = ([f"fpn_p{l}"], None)
    m["rpn.shared"] = (["rpn_conv_shared"], None)
    m["rpn.head"] = (["rpn_class_raw", "rpn_bbox_pred"], None)           # concatenated along the output channels: 6 + 12
    m["cls.conv1"] = (["mrcnn_class_conv1"], "mrcnn_class_bn1")
    m["cls.conv2"] = (["mrcnn_class_conv2"], "mrcnn_class_bn2")
    m["cls.fc"] = (["mrcnn_class_logits", "mrcnn_bbox_fc"], None)        # 81 logits | 324 deltas
    for i in range(1, 5):
        m[f"mask.conv{i}"] = ([f"mrcnn_mask_conv{i}"], f"mrcnn_mask_bn{i}")
    m["mask.deconv"] = (["mrcnn_mask_deconv"], None)
    m["mask.final"] = (["mrcnn_mask"], None)
    return m


def _kernel_hwio(L):
    """Core ML layer -> Keras-layout kernel (HWIO; Conv2DTranspose [kh, kw, out, in]) + bias."""
    w = L["weights"]
    if w is None:
        raise ValueError(f"layer {L['name']} has no weights")
    if L["type"] == "innerProduct":
        k = w.T.reshape(1, 1, L["cin"], L["cout"])
    else:
        k = np.transpose(w, (2, 3, 1, 0))        # [Cout,Cin,kh,kw] -> HWIO; deconv [Cin,Cout,kh,kw] -> [kh,kw,Cout,Cin]
    cout = L["cout"]
    b = L["bias"] if L.get("bias") is not None else np.zeros(cout, np.float32)
    return np.ascontiguousarray(k, np.float32), np.asarray(b, np.float32)


def params_from_mlmodels(models, architecture=101, num_classes=81, pool_classifier=7):
    """models: parsed (read_mlmodel) MaskRCNN / Classifier / Mask models, any order or merged ->
    reference-layout parameters for weights.fold() plus {"mean_rgb", "custom_layers"} read from the main model."""
    by_name, customs, pre = {}, {}, {}
    for mdl in models:
        for L in mdl["layers"]:
            by_name[L["name"]] = L
            if L["type"] == "custom":
                customs[L["class_name"]] = L["parameters"]
        if mdl["preprocessing"]:
            pre = mdl["preprocessing"]

    def find(name):
        # the Keras converter keeps the Keras layer name; TimeDistributed wrappers may prefix / suffix it
        if name in by_name:
            return by_name[name]
        hits = [L for n, L in by_name.items() if name in n and L["type"] in ("convolution", "deconvolution", "innerProduct", "batchnorm")]
        if len(hits) == 1:
            return hits[0]
        raise KeyError(f"layer '{name}' not found in the Core ML models ({len(hits)} candidates)")

    names = keras_names(architecture)
    params = {}
    for name, which, kind, kh, kw, cin, cout, has_bn, relu in W.layer_table(architecture, num_classes, pool_classifier):
        convs, bn = names[name]
        ks, bs = zip(*[_kernel_hwio(find(c)) for c in convs])
        kernel, bias = np.concatenate(ks, axis=-1 if kind != "deconv" else 2), np.concatenate(bs)
        want = (kh, kw, cout, cin) if kind == "deconv" else (kh, kw, cin, cout)
        if kernel.shape != want:
            raise ValueError(f"{name}: kernel shape {kernel.shape}, expected {want}")
        p = {"kernel": kernel, "bias": bias}
        if bn is not None:
            try:
                B = find(bn)
                p["bn"] = (B["gamma"], B["beta"], B["mean"], B["variance"])
                p["bn_eps"] = B["epsilon"]
            except KeyError:
                pass                              # the converter may have fused the batch norm into the convolution
        params[name] = p
    extra = {"custom_layers": customs}
    if pre:
        extra["mean_rgb"] = (-pre["redBias"], -pre["greenBias"], -pre["blueBias"])      # Conversion/task.py:73-75
    return params, extra


def mlmodels_from_params(params, architecture=101, num_classes=81, pool_classifier=7, half=True, custom_layers=None,
                         mean_rgb=(123.7, 116.8, 103.9)):
    """The inverse: reference-layout parameters -> [MaskRCNN, Classifier, Mask] .mlmodel bytes holding the weight-carrying
    layers under their Keras names (+ the custom layers' parameter maps): what the tests round-trip, and an export path."""
    names = keras_names(architecture)
    layers = {W.MAIN: [], W.CLASSIFIER: [], W.MASK: []}
    for name, which, kind, kh, kw, cin, cout, has_bn, relu in W.layer_table(architecture, num_classes, pool_classifier):
        convs, bn = names[name]
        p = params[name]
        splits = {"rpn.head": [6, 12], "cls.fc": [num_classes, num_classes * 4]}.get(name, [cout])
        o0 = 0
        for cname, n in zip(convs, splits):
            if kind == "deconv":
                k = p["kernel"][:, :, o0:o0 + n, :]
                wts = np.transpose(k, (3, 2, 0, 1))                 # [kh,kw,out,in] -> [Cin,Cout,kh,kw]
                layers[which].append({"name": cname, "type": "deconvolution", "inputs": [cname + "_in"], "outputs": [cname + "_out"],
                                      "cout": n, "kernel_channels": cin, "kernel_size": (kh, kw), "stride": (2, 2), "padding": "valid",
                                      "weights": wts, "bias": p["bias"][o0:o0 + n]})
            elif name == "cls.fc":
                k = p["kernel"][0, 0, :, o0:o0 + n]
                layers[which].append({"name": cname, "type": "innerProduct", "inputs": [cname + "_in"], "outputs": [cname + "_out"],
                                      "cin": cin, "cout": n, "weights": k.T, "bias": p["bias"][o0:o0 + n]})
            else:
                k = p["kernel"][..., o0:o0 + n]
                layers[which].append({"name": cname, "type": "convolution", "inputs": [cname + "_in"], "outputs": [cname + "_out"],
                                      "cout": n, "kernel_channels": cin, "kernel_size": (kh, kw), "stride": (1, 1),
                                      "padding": "same" if kh > 1 and not name.startswith("cls.") else "valid",
                                      "weights": np.transpose(k, (3, 2, 0, 1)), "bias": p["bias"][o0:o0 + n]})
            o0 += n
        if bn is not None and "bn" in p:
            g, b, mu, var = p["bn"]
            layers[which].append({"name": bn, "type": "batchnorm", "inputs": [bn + "_in"], "outputs": [bn + "_out"], "channels": cout,
                                  "epsilon": float(p.get("bn_eps", W.BN_EPS)), "gamma": g, "beta": b, "mean": mu, "variance": var})
    for cls, prm in (custom_layers or {}).items():
        layers[W.MAIN].append({"name": cls.lower(), "type": "custom", "inputs": [], "outputs": [], "class_name": cls, "parameters": prm})
    pre = {"redBias": -mean_rgb[0], "greenBias": -mean_rgb[1], "blueBias": -mean_rgb[2], "channelScale": 1.0}
    return [write_mlmodel(layers[W.MAIN], pre, half), write_mlmodel(layers[W.CLASSIFIER], None, half), write_mlmodel(layers[W.MASK], None, half)]


def config_from_custom_layers(customs, config):
    """Copies the parameters Core ML bakes into the custom layers (ProposalLayer.swift:65-91, DetectionLayer.swift:63-88,
    PyramidROIAlignLayer.swift:48-59) into a MaskRCNNConfig."""
    p = customs.get("ProposalLayer", {})
    if "preNMSMaxProposals" in p:
        config.preNMSMaxProposals = int(p["preNMSMaxProposals"])
    if "maxProposals" in p:
        config.maxProposals = int(p["maxProposals"])
    d = customs.get("DetectionLayer", {})
    if "maxDetections" in d:
        config.maxDetections = int(d["maxDetections"])
    r = customs.get("PyramidROIAlignLayer", {})
    if "imageWidth" in r and "imageHeight" in r:
        config.imageShape = (int(r["imageHeight"]), int(r["imageWidth"]), 3)
    return config


def import_products(main, classifier, mask, architecture=101, num_classes=81, pool_classifier=7):
    """Three .mlmodel files (paths or bytes) -> (folded weights, [main, classifier, mask] MRCNNW1 blobs, extra)."""
    models = [read_mlmodel(x) for x in (main, classifier, mask)]
    params, extra = params_from_mlmodels(models, architecture, num_classes, pool_classifier)
    folded = W.fold(params, architecture, num_classes, pool_classifier)
    blobs = [W.pack_blob(w, W.device_tensors(folded, w, architecture, num_classes, pool_classifier)) for w in (W.MAIN, W.CLASSIFIER, W.MASK)]
    return folded, blobs, extra
