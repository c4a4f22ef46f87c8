"""`Results` protobuf writer / reader: the wire format the reference hands to its COCO evaluation task
(Sources/maskrcnn/results.pb.swift; filled at EvaluateCommand.swift:180-198 and :203-248, parsed by
Python/COCOEval/task.py:93-96).  SURVEY.md section 8 f3.  Hand-rolled proto3 encoding (no generated code):

    message Results { repeated Result results = 1; }
    message Result  { ImageInfo imageInfo = 1; repeated Detection detections = 2;
        message Origin { double x = 1; double y = 2; }
        message Size   { double width = 1; double height = 2; }
        message Rect   { Origin origin = 1; Size size = 2; }
        message ImageInfo { string datasetId = 1; string id = 2; int32 width = 3; int32 height = 4; }
        message Detection { double probability = 1; int32 classId = 2; string classLabel = 3; Rect boundingBox = 4; } }

proto3 semantics: scalar fields equal to their default (0, 0.0, "") are not emitted.
"""
import struct


def _varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _double(field, v):
    return b"" if v == 0.0 else _varint((field << 3) | 1) + struct.pack("<d", float(v))


def _int32(field, v):
    return b"" if v == 0 else _varint((field << 3) | 0) + _varint(int(v))


def _string(field, s):
    b = s.encode()
    return b"" if not b else _varint((field << 3) | 2) + _varint(len(b)) + b


def _message(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_detection(probability, class_id, class_label, x, y, width, height):
    rect = _message(1, _double(1, x) + _double(2, y)) + _message(2, _double(1, width) + _double(2, height))
    return _double(1, probability) + _int32(2, class_id) + _string(3, class_label) + _message(4, rect)


def encode_result(dataset_id, image_id, width, height, detections):
    """detections: iterable of (probability, class_id, class_label, x, y, w, h)."""
    info = _string(1, dataset_id) + _string(2, str(image_id)) + _int32(3, width) + _int32(4, height)
    return _message(1, info) + b"".join(_message(2, encode_detection(*d)) for d in detections)


def encode_results(results):
    """results: iterable of already encoded Result payloads."""
    return b"".join(_message(1, r) for r in results)


def result_from_detections(dataset_id, image_id, width, height, detections, class_label="test"):
    """The reference's conversion (EvaluateCommand.swift:203-248): rows of the (D,6) `detections` output with
    Double(score) > 0.7 become Detection messages with boundingBox = (x1, y1, x2-x1, y2-y1) in Double and the
    hard-coded label "test" (:221)."""
    dets = []
    for row in detections:
        y1, x1, y2, x2, cls, score = [float(v) for v in row]
        if score > 0.7:
            dets.append((score, int(cls), class_label, x1, y1, x2 - x1, y2 - y1))
    return encode_result(dataset_id, image_id, width, height, dets)


# ---- minimal decoder (tests; also lets a Python consumer read the file without generated code) ----
# Every malformed input raises ValueError (truncation, wrong wire type for a field, bad UTF-8, over-long varints).
def _read_varint(buf, pos):
    shift = result = 0
    while True:
        if pos >= len(buf) or shift > 63:
            raise ValueError("Results: truncated or over-long varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf):
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            if pos + 8 > len(buf):
                raise ValueError("Results: truncated double")
            v = struct.unpack_from("<d", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            if pos + n > len(buf):
                raise ValueError("Results: truncated length-delimited field")
            v = bytes(buf[pos:pos + n])
            pos += n
        else:
            raise ValueError(f"Results: unsupported wire type {wt}")
        yield field, wt, v


def _want(wt, expected, what):
    if wt != expected:
        raise ValueError(f"Results: field {what} has wire type {wt}, expected {expected}")


def _text(b):
    try:
        return b.decode("utf-8")
    except UnicodeDecodeError:
        raise ValueError("Results: string field is not UTF-8") from None


def _to_int32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def decode_results(buf):
    """Unknown fields are skipped, as protobuf parsers do."""
    out = []
    for f, wt, payload in _fields(bytes(buf)):
        if f != 1:
            continue
        _want(wt, 2, "Results.results")
        res = {"imageInfo": {"datasetId": "", "id": "", "width": 0, "height": 0}, "detections": []}
        for f2, wt2, v in _fields(payload):
            if f2 == 1:
                _want(wt2, 2, "Result.imageInfo")
                for f3, wt3, w in _fields(v):
                    if f3 in (1, 2):
                        _want(wt3, 2, "ImageInfo string")
                        res["imageInfo"]["datasetId" if f3 == 1 else "id"] = _text(w)
                    elif f3 in (3, 4):
                        _want(wt3, 0, "ImageInfo int32")
                        res["imageInfo"]["width" if f3 == 3 else "height"] = _to_int32(w)
            elif f2 == 2:
                _want(wt2, 2, "Result.detections")
                d = {"probability": 0.0, "classId": 0, "classLabel": "", "boundingBox": {"x": 0.0, "y": 0.0, "width": 0.0, "height": 0.0}}
                for f3, wt3, w in _fields(v):
                    if f3 == 1:
                        _want(wt3, 1, "Detection.probability")
                        d["probability"] = w
                    elif f3 == 2:
                        _want(wt3, 0, "Detection.classId")
                        d["classId"] = _to_int32(w)
                    elif f3 == 3:
                        _want(wt3, 2, "Detection.classLabel")
                        d["classLabel"] = _text(w)
                    elif f3 == 4:
                        _want(wt3, 2, "Detection.boundingBox")
                        for f4, wt4, r in _fields(w):
                            if f4 not in (1, 2):
                                continue
                            _want(wt4, 2, "Rect.origin / Rect.size")
                            names = ("x", "y") if f4 == 1 else ("width", "height")
                            for f5, wt5, val in _fields(r):
                                if f5 in (1, 2):
                                    _want(wt5, 1, "Origin / Size double")
                                    d["boundingBox"][names[f5 - 1]] = val
                res["detections"].append(d)
        out.append(res)
    return out
