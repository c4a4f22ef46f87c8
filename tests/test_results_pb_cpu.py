"""The `Results` protobuf the reference's evaluation task consumes (results.pb.swift): hand-rolled encoder checked
against google.protobuf's dynamic messages built from the same schema, and against the reference's filtering rule."""
import numpy as np
import pytest


def _dynamic_classes():
    pb = pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="results.proto", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto
    res = fd.message_type.add(name="Result")

    def msg(parent, name, fields):
        m = parent.nested_type.add(name=name)
        for i, (fname, ftype, tname) in enumerate(fields, 1):
            f = m.field.add(name=fname, number=i, type=ftype, label=T.LABEL_OPTIONAL)
            if tname:
                f.type_name = tname
        return m
    msg(res, "Origin", [("x", T.TYPE_DOUBLE, None), ("y", T.TYPE_DOUBLE, None)])
    msg(res, "Size", [("width", T.TYPE_DOUBLE, None), ("height", T.TYPE_DOUBLE, None)])
    msg(res, "Rect", [("origin", T.TYPE_MESSAGE, ".Result.Origin"), ("size", T.TYPE_MESSAGE, ".Result.Size")])
    msg(res, "ImageInfo", [("datasetId", T.TYPE_STRING, None), ("id", T.TYPE_STRING, None), ("width", T.TYPE_INT32, None), ("height", T.TYPE_INT32, None)])
    msg(res, "Detection", [("probability", T.TYPE_DOUBLE, None), ("classId", T.TYPE_INT32, None), ("classLabel", T.TYPE_STRING, None),
                           ("boundingBox", T.TYPE_MESSAGE, ".Result.Rect")])
    res.field.add(name="imageInfo", number=1, type=T.TYPE_MESSAGE, label=T.LABEL_OPTIONAL, type_name=".Result.ImageInfo")
    res.field.add(name="detections", number=2, type=T.TYPE_MESSAGE, label=T.LABEL_REPEATED, type_name=".Result.Detection")
    rs = fd.message_type.add(name="Results")
    rs.field.add(name="results", number=1, type=T.TYPE_MESSAGE, label=T.LABEL_REPEATED, type_name=".Result")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("Results"))


def test_results_wire_format_matches_google_protobuf(pkg):
    Results = _dynamic_classes()
    rng = np.random.default_rng(0)
    det = np.zeros((100, 6), np.float32)
    det[:20, :4] = np.sort(rng.uniform(size=(20, 4)).astype(np.float32), axis=1)
    det[:20, 4] = rng.integers(1, 81, 20)
    det[:20, 5] = rng.uniform(0.6, 1.0, 20)
    det[3, 5] = np.float32(0.7)                       # Double(Float(0.7)) < 0.7: dropped (EvaluateCommand.swift:217)
    blob = pkg.results_pb.encode_results([pkg.results_pb.result_from_detections("coco", 139, 640, 426, det),
                                          pkg.results_pb.result_from_detections("coco", 285, 0, 0, det[:0])])
    msg = Results()
    msg.ParseFromString(blob)
    assert len(msg.results) == 2 and msg.results[0].imageInfo.id == "139" and msg.results[0].imageInfo.width == 640
    keep = [r for r in det if float(r[5]) > 0.7]
    assert len(msg.results[0].detections) == len(keep) and 3 not in [i for i, r in enumerate(det) if float(r[5]) > 0.7]
    for d, r in zip(msg.results[0].detections, keep):
        assert d.probability == float(r[5]) and d.classId == int(r[4]) and d.classLabel == "test"
        assert d.boundingBox.origin.x == float(r[1]) and d.boundingBox.origin.y == float(r[0])
        assert d.boundingBox.size.width == float(r[3]) - float(r[1]) and d.boundingBox.size.height == float(r[2]) - float(r[0])
    assert msg.SerializeToString(deterministic=True) == blob           # byte-identical to the canonical encoder
    # and our own decoder round-trips
    back = pkg.results_pb.decode_results(blob)
    assert back[0]["imageInfo"] == {"datasetId": "coco", "id": "139", "width": 640, "height": 426}
    assert len(back[0]["detections"]) == len(keep) and back[1]["detections"] == []


def test_decoder_rejects_damaged_messages_with_valueerror(pkg):
    rng = np.random.default_rng(0)
    det = np.zeros((100, 6), np.float32)
    det[:20, :4] = np.sort(rng.uniform(size=(20, 4)).astype(np.float32), axis=1)
    det[:20, 4] = rng.integers(1, 81, 20)
    det[:20, 5] = rng.uniform(0.71, 1.0, 20)
    good = pkg.results_pb.encode_results([pkg.results_pb.result_from_detections("coco", 139, 640, 426, det)])
    assert len(pkg.results_pb.decode_results(good)[0]["detections"]) == 20
    ok = 0
    for _ in range(3000):                               # truncations and byte flips: ValueError or a clean parse, nothing else
        b = bytearray(good)
        if rng.random() < 0.3:
            b = b[:int(rng.integers(0, len(b)))]
        else:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        try:
            pkg.results_pb.decode_results(bytes(b))
            ok += 1
        except ValueError:
            pass
    assert 0 < ok < 3000
