"""ctypes wrapper around oracle/liboracle.so (the CPU restatement, oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of oracle.c.  Importable from
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
PARITY UNPINNED: the reference has no golden vectors and cannot run here.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
            build()
        _lib = C.CDLL(_LIB)
        _lib.orc_iou.restype = C.c_float
        for f in ("orc_nms", "orc_proposal", "orc_detection", "orc_detections_decode"):
            getattr(_lib, f).restype = C.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def iou(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_iou(_p(a), _p(b)))


def apply_box_deltas(boxes, deltas):
    b, d = _f32(boxes).copy(), _f32(deltas)
    lib().orc_apply_box_deltas(_p(b), _p(d), C.c_int(b.shape[0]))
    return b


def clip(boxes):
    b = _f32(boxes).copy()
    lib().orc_clip(_p(b), C.c_int(b.size))
    return b


def nms(boxes, indices, iou_thr, max_keep):
    b = _f32(boxes)
    idx = np.ascontiguousarray(indices, dtype=np.int32)
    sel = np.zeros(max(max_keep, 1), dtype=np.int32)
    n = lib().orc_nms(_p(b), _p(idx), C.c_int(idx.size), C.c_float(iou_thr), C.c_int(max_keep), _p(sel))
    return sel[:n].copy()


def argsort_desc(keys):
    k = _f32(keys)
    out = np.zeros(k.size, dtype=np.uint32)
    lib().orc_argsort_desc(_p(k), C.c_int(k.size), _p(out))
    return out


def proposal(probs, deltas, anchors, std=(0.1, 0.1, 0.2, 0.2), pre_nms=6000, max_proposals=1000, iou_thr=0.7,
             return_sorted_boxes=False):
    """One image: returns (rois (max,4), keep_anchor (max,), count)."""
    probs, deltas, anchors = _f32(probs), _f32(deltas), _f32(anchors)
    n = probs.shape[0]
    rois = np.zeros((max_proposals, 4), dtype=np.float32)
    keep = np.zeros(max_proposals, dtype=np.int32)
    sd = _f32(std)
    sb = np.zeros((min(n, pre_nms), 4), dtype=np.float32) if return_sorted_boxes else None
    cnt = lib().orc_proposal(_p(probs), _p(deltas), _p(anchors), C.c_int(n), _p(sd), C.c_int(pre_nms),
                             C.c_int(max_proposals), C.c_float(np.float32(iou_thr)), _p(rois), _p(keep), _p(sb))
    if return_sorted_boxes:
        return rois, keep, cnt, sb
    return rois, keep, cnt


def roi_levels(rois, image_w=1024, image_h=1024):
    r = _f32(rois)
    out = np.zeros(r.shape[0], dtype=np.int32)
    lib().orc_roi_levels(_p(r), C.c_int(r.shape[1]), C.c_int(r.shape[0]), C.c_double(image_w), C.c_double(image_h), _p(out))
    return out


def pyramid_roialign(rois, fmaps, pool, image_w=1024, image_h=1024):
    """One image: rois (R,4|6), fmaps = 4 CHW f32 arrays -> (out (R,C,P,P), levels (R,))."""
    r = _f32(rois)
    fm = [_f32(f) for f in fmaps]
    c = fm[0].shape[0]
    hw = np.array([d for f in fm for d in f.shape[1:]], dtype=np.int32)
    out = np.empty((r.shape[0], c, pool, pool), dtype=np.float32)
    lv = np.zeros(r.shape[0], dtype=np.int32)
    lib().orc_pyramid_roialign(_p(r), C.c_int(r.shape[1]), C.c_int(r.shape[0]), _p(fm[0]), _p(fm[1]), _p(fm[2]),
                               _p(fm[3]), _p(hw), C.c_int(c), C.c_int(pool), C.c_double(image_w),
                               C.c_double(image_h), _p(out), _p(lv))
    return out, lv


def pyramid_roialign_nhwc_f16(rois, fmaps_hwc_f16, pool, image_w=1024, image_h=1024):
    """One image: fmaps = 4 HWC float16 arrays -> (out (R,P,P,C) float16, levels)."""
    r = _f32(rois)
    fm = [np.ascontiguousarray(f, dtype=np.float16) for f in fmaps_hwc_f16]
    c = fm[0].shape[2]
    hw = np.array([d for f in fm for d in f.shape[:2]], dtype=np.int32)
    out = np.empty((r.shape[0], pool, pool, c), dtype=np.float16)
    lv = np.zeros(r.shape[0], dtype=np.int32)
    lib().orc_pyramid_roialign_nhwc_f16(_p(r), C.c_int(r.shape[1]), C.c_int(r.shape[0]), _p(fm[0]), _p(fm[1]),
                                        _p(fm[2]), _p(fm[3]), _p(hw), C.c_int(c), C.c_int(pool),
                                        C.c_double(image_w), C.c_double(image_h), _p(out), _p(lv))
    return out, lv


def classifier_select(probs, bbox):
    p, b = _f32(probs), _f32(bbox)
    out = np.empty((p.shape[0], 6), dtype=np.float32)
    lib().orc_classifier_select(_p(p), _p(b), C.c_int(p.shape[0]), C.c_int(p.shape[1]), _p(out))
    return out


def detection(rois, cls, std=(0.1, 0.1, 0.2, 0.2), max_det=100, score_thr=0.7, iou_thr=0.3):
    """One image: returns (detections (max_det,6), keep_roi (max_det,), count)."""
    r, c = _f32(rois), _f32(cls)
    out = np.zeros((max_det, 6), dtype=np.float32)
    keep = np.zeros(max_det, dtype=np.int32)
    sd = _f32(std)
    cnt = lib().orc_detection(_p(r), _p(c), C.c_int(r.shape[0]), _p(sd), C.c_int(max_det),
                              C.c_float(np.float32(score_thr)), C.c_float(np.float32(iou_thr)), _p(out), _p(keep))
    return out, keep, cnt


def mask_select(masks_all, valid, detections):
    m, d = _f32(masks_all), _f32(detections)
    v = np.ascontiguousarray(valid, dtype=np.int32)
    dn, ncls, s = m.shape[0], m.shape[1], m.shape[2]
    out = np.empty((dn, s, s), dtype=np.float32)
    lib().orc_mask_select(_p(m), _p(v), _p(d), C.c_int(dn), C.c_int(ncls), C.c_int(s), _p(out))
    return out


def detections_decode(det, masks=None):
    d = _f32(det)
    dn = d.shape[0]
    s = masks.shape[-1] if masks is not None else 28
    m = _f32(masks) if masks is not None else None
    idx = np.zeros(dn, dtype=np.int32)
    bbox = np.zeros((dn, 4), dtype=np.float64)
    cls = np.zeros(dn, dtype=np.int32)
    score = np.zeros(dn, dtype=np.float64)
    mu8 = np.zeros((dn, s * s), dtype=np.uint8) if masks is not None else None
    n = lib().orc_detections_decode(_p(d), _p(m), C.c_int(dn), C.c_int(s), _p(idx), _p(bbox), _p(cls), _p(score), _p(mu8))
    return n, idx, bbox, cls, score, mu8


def letterbox(src, dst_h=1024, dst_w=1024):
    s = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.empty((dst_h, dst_w, 3), dtype=np.uint8)
    lib().orc_letterbox(_p(s), C.c_int(s.shape[0]), C.c_int(s.shape[1]), C.c_int(dst_h), C.c_int(dst_w), _p(out))
    return out
