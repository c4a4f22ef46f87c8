"""the .mlmodel importer (mask-rcnn-coreml_b200/mlmodel.py).  No artefact of the reference is reachable
offline (PARITY UNPINNED); these tests pin what can be pinned here: the protobuf wire encoding against google.protobuf,
and the round trip reference-layout parameters -> three .mlmodel files -> importer -> MRCNNW1 blobs."""
import numpy as np
import pytest


def _params16(pkg, arch):
    """Synthetic parameters rounded to fp16, as the converter stores them (Conversion/task.py:90,102,114)."""
    params = pkg.weights.synthetic(arch)
    for p in params.values():
        p["kernel"] = p["kernel"].astype(np.float16).astype(np.float32)
        p["bias"] = p["bias"].astype(np.float16).astype(np.float32)
        if "bn" in p:
            p["bn"] = tuple(x.astype(np.float16).astype(np.float32) for x in p["bn"])
            p["bn_eps"] = float(np.float32(pkg.weights.BN_EPS))       # the file stores epsilon as a 32-bit float
    return params


@pytest.fixture(scope="module")
def exported(pkg):
    params = _params16(pkg, 50)
    customs = {"ProposalLayer": {"preNMSMaxProposals": 6000, "maxProposals": 1000, "nmsIOUThreshold": 0.7, "bboxStdDev_count": 4},
               "DetectionLayer": {"maxDetections": 100, "scoreThreshold": 0.7}, "PyramidROIAlignLayer": {"poolSize": 7, "imageWidth": 1024, "imageHeight": 1024}}
    files = pkg.mlmodel.mlmodels_from_params(params, 50, custom_layers=customs)
    return params, customs, files


def test_round_trip_blobs_identical(pkg, exported):
    params, customs, files = exported
    folded, blobs, extra = pkg.mlmodel.import_products(*files, architecture=50)
    want = pkg.weights.fold(params, 50)
    assert set(folded) == set(want)
    for k in want:
        np.testing.assert_array_equal(folded[k][0], want[k][0], err_msg=k)
        np.testing.assert_array_equal(folded[k][1], want[k][1], err_msg=k)
    for which, blob in enumerate(blobs):
        assert blob == pkg.weights.pack_blob(which, pkg.weights.device_tensors(want, which, 50))
    assert extra["custom_layers"] == customs
    np.testing.assert_allclose(extra["mean_rgb"], (123.7, 116.8, 103.9), rtol=1e-6)       # Conversion/task.py:73-75
    cfg = pkg.mlmodel.config_from_custom_layers(extra["custom_layers"], pkg.MaskRCNNConfig())
    assert (cfg.preNMSMaxProposals, cfg.maxProposals, cfg.maxDetections, cfg.imageShape) == (6000, 1000, 100, (1024, 1024, 3))


def test_float32_weights_and_fused_batchnorm(pkg):
    """floatValue (packed f32) weights, and a model whose batch norms were fused by the converter (no batchnorm layers)."""
    params = _params16(pkg, 50)
    for p in params.values():
        p.pop("bn", None)
    files = pkg.mlmodel.mlmodels_from_params(params, 50, half=False)
    folded, _, _ = pkg.mlmodel.import_products(*files, architecture=50)
    want = pkg.weights.fold(params, 50)
    for k in want:
        np.testing.assert_array_equal(folded[k][0], want[k][0], err_msg=k)


def test_layer_names_follow_matterport(pkg):
    n = pkg.mlmodel.keras_names(101)
    assert n["res4.0.2a"] == (["res4a_branch2a"], "bn4a_branch2a")
    assert n["res4.22.2c"] == (["res4w_branch2c"], "bn4w_branch2c")          # 22 identity blocks 'b'..'w'
    assert n["res2.0.1"] == (["res2a_branch1"], "bn2a_branch1")
    assert n["rpn.head"][0] == ["rpn_class_raw", "rpn_bbox_pred"] and n["cls.fc"][0] == ["mrcnn_class_logits", "mrcnn_bbox_fc"]
    assert set(n) == {t[0] for t in pkg.weights.layer_table(101)}


def test_missing_layer_and_bad_file_fail_loudly(pkg, exported):
    _, _, files = exported
    with pytest.raises(ValueError, match="not a neural-network"):
        pkg.mlmodel.read_mlmodel(pkg.mlmodel.enc_varint(1, 2))
    with pytest.raises(KeyError, match="mrcnn_mask"):
        pkg.mlmodel.import_products(files[0], files[1], pkg.mlmodel.write_mlmodel([]), architecture=50)
    with pytest.raises(ValueError):
        pkg.mlmodel.parse_message(files[2][:1000])                                # truncated


def test_wire_format_matches_google_protobuf(pkg):
    """The hand-written encoder / decoder against google.protobuf on the same (sub-)schema, built at run time."""
    pb = pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    T = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="coreml_subset.proto", package="cm", syntax="proto3")

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = ".cm." + tname
    O, R = T.LABEL_OPTIONAL, T.LABEL_REPEATED
    msg("WeightParams", [("floatValue", 1, T.TYPE_FLOAT, R, None), ("float16Value", 2, T.TYPE_BYTES, O, None)])
    msg("Valid", [])
    msg("Conv", [("outputChannels", 1, T.TYPE_UINT64, O, None), ("kernelChannels", 2, T.TYPE_UINT64, O, None), ("nGroups", 10, T.TYPE_UINT64, O, None),
                 ("kernelSize", 20, T.TYPE_UINT64, R, None), ("stride", 30, T.TYPE_UINT64, R, None), ("valid", 50, T.TYPE_MESSAGE, O, "Valid"),
                 ("same", 51, T.TYPE_MESSAGE, O, "Valid"), ("isDeconvolution", 60, T.TYPE_BOOL, O, None), ("hasBias", 70, T.TYPE_BOOL, O, None),
                 ("weights", 90, T.TYPE_MESSAGE, O, "WeightParams"), ("bias", 91, T.TYPE_MESSAGE, O, "WeightParams")])
    msg("BN", [("channels", 1, T.TYPE_UINT64, O, None), ("epsilon", 10, T.TYPE_FLOAT, O, None), ("gamma", 15, T.TYPE_MESSAGE, O, "WeightParams"),
               ("beta", 16, T.TYPE_MESSAGE, O, "WeightParams"), ("mean", 17, T.TYPE_MESSAGE, O, "WeightParams"), ("variance", 18, T.TYPE_MESSAGE, O, "WeightParams")])
    msg("Layer", [("name", 1, T.TYPE_STRING, O, None), ("input", 2, T.TYPE_STRING, R, None), ("output", 3, T.TYPE_STRING, R, None),
                  ("convolution", 100, T.TYPE_MESSAGE, O, "Conv"), ("batchnorm", 160, T.TYPE_MESSAGE, O, "BN")])
    msg("NN", [("layers", 1, T.TYPE_MESSAGE, R, "Layer")])
    msg("Model", [("specificationVersion", 1, T.TYPE_INT32, O, None), ("neuralNetwork", 500, T.TYPE_MESSAGE, O, "NN")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    Model = message_factory.GetMessageClass(pool.FindMessageTypeByName("cm.Model"))
    rng = np.random.default_rng(3)
    w = rng.standard_normal((8, 4, 3, 3)).astype(np.float16).astype(np.float32)
    b = rng.standard_normal(8).astype(np.float16).astype(np.float32)
    var = (b * b).astype(np.float16).astype(np.float32)
    layers = [{"name": "c", "type": "convolution", "inputs": ["x"], "outputs": ["y"], "cout": 8, "kernel_channels": 4, "kernel_size": (3, 3),
               "stride": (1, 1), "padding": "same", "weights": w, "bias": b},
              {"name": "bn", "type": "batchnorm", "inputs": ["y"], "outputs": ["z"], "channels": 8, "epsilon": 1e-3, "gamma": b, "beta": b, "mean": b, "variance": var}]
    # ours -> google.protobuf
    for half in (True, False):
        data = pkg.mlmodel.write_mlmodel(layers, half=half)
        m = Model()
        m.ParseFromString(data)
        assert m.specificationVersion == 2 and len(m.neuralNetwork.layers) == 2
        c = m.neuralNetwork.layers[0].convolution
        assert (c.outputChannels, c.kernelChannels, list(c.kernelSize), list(c.stride), c.hasBias, c.HasField("same")) == (8, 4, [3, 3], [1, 1], True, True)
        got = np.frombuffer(c.weights.float16Value, "<f2").astype(np.float32) if half else np.array(c.weights.floatValue, np.float32)
        np.testing.assert_array_equal(got, w.ravel())
        assert abs(m.neuralNetwork.layers[1].batchnorm.epsilon - 1e-3) < 1e-9
        # google.protobuf -> ours (its serialisation uses packed repeated scalars)
        back = pkg.mlmodel.read_mlmodel(m.SerializeToString())
        np.testing.assert_array_equal(back["layers"][0]["weights"], w)
        np.testing.assert_array_equal(back["layers"][0]["bias"], b)
        assert back["layers"][0]["kernel_size"] == (3, 3) and back["layers"][0]["padding"] == "same"
        np.testing.assert_array_equal(back["layers"][1]["variance"], var)


def test_parser_rejects_garbage_cleanly(pkg):
    """Random bytes and byte mutations of a small model that exercises every layer kind either parse or raise
    ValueError / KeyError -- never another exception type (wrong wire types under known field numbers, truncated
    scalars, odd-length weight blobs ...), never a hang."""
    rng = np.random.default_rng(11)
    w = rng.standard_normal((8, 4, 3, 3)).astype(np.float32)
    b = rng.standard_normal(8).astype(np.float32)
    layers = [{"name": "c", "type": "convolution", "inputs": ["x"], "outputs": ["y"], "cout": 8, "kernel_channels": 4, "kernel_size": (3, 3),
               "stride": (1, 1), "padding": "same", "weights": w, "bias": b},
              {"name": "d", "type": "deconvolution", "inputs": ["x"], "outputs": ["y"], "cout": 8, "kernel_channels": 4, "kernel_size": (2, 2),
               "stride": (2, 2), "padding": "valid", "weights": rng.standard_normal((4, 8, 2, 2)).astype(np.float32), "bias": b},
              {"name": "bn", "type": "batchnorm", "inputs": ["y"], "outputs": ["z"], "channels": 8, "epsilon": 1e-3, "gamma": b, "beta": b, "mean": b, "variance": b * b},
              {"name": "fc", "type": "innerProduct", "inputs": ["z"], "outputs": ["o"], "cin": 4, "cout": 8, "weights": rng.standard_normal((8, 4)).astype(np.float32), "bias": b},
              {"name": "p", "type": "custom", "inputs": [], "outputs": [], "class_name": "ProposalLayer",
               "parameters": {"maxProposals": 1000, "thr": 0.7, "s": "x", "flag": True}}]
    parsed = 0
    for half in (True, False):
        src = pkg.mlmodel.write_mlmodel(layers, {"redBias": -1.0}, half)
        back = pkg.mlmodel.read_mlmodel(src)
        assert [L["type"] for L in back["layers"]] == [L["type"] for L in layers]
        assert back["layers"][4]["parameters"] == layers[4]["parameters"] and back["preprocessing"]["redBias"] == -1.0
        cases = [bytes(rng.integers(0, 256, int(n), dtype=np.uint8)) for n in rng.integers(1, 200, 50)]
        cases += [src[:int(k)] for k in rng.integers(1, len(src), 50)]
        for _ in range(3000):
            m = bytearray(src)
            for pos in rng.integers(0, len(m), int(rng.integers(1, 6))):
                m[int(pos)] = int(rng.integers(0, 256))
            cases.append(bytes(m))
        for data in cases:
            try:
                pkg.mlmodel.read_mlmodel(data)
                parsed += 1
            except (ValueError, KeyError):
                pass
    assert parsed > 1000          # most single-byte mutations land in weight payloads and still parse


def test_import_cli_writes_products(pkg, exported, tmp_path):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = ("MaskRCNN.mlmodel", "Classifier.mlmodel", "Mask.mlmodel")
    for n, data in zip(names, exported[2]):
        (tmp_path / n).write_bytes(data)
    out = tmp_path / "products"
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "import_mlmodel.py"), "--main", str(tmp_path / names[0]),
                           "--classifier", str(tmp_path / names[1]), "--mask", str(tmp_path / names[2]), "--architecture", "50",
                           "--out", str(out), "--anchors"], stdout=subprocess.DEVNULL)
    want = pkg.weights.fold(exported[0], 50)
    for which, n in enumerate(("MaskRCNN", "Classifier", "Mask")):
        blob = (out / (n + ".mrcnnw")).read_bytes()
        assert blob == pkg.weights.pack_blob(which, pkg.weights.device_tensors(want, which, 50))
    anchors = np.fromfile(out / "anchors.bin", np.float32).reshape(-1, 4)
    assert anchors.shape[0] == 261888 and (anchors == pkg.synth.generate_anchors(1024, 1024)).all()      # Conversion/task.py:176
