"""Deterministic synthetic inputs for the hot path.

No datasets or trained weights are reachable offline, so tests and bench.py use
seeded inputs shaped like the real ones: COCO-like clustered RPN outputs (so NMS
suppresses most boxes yet still reaches max_proposals), N(0,1) feature maps and
classifier outputs with a realistic foreground fraction.
Seed convention: numpy.random.default_rng(20260 + image_index).
"""
import math

import numpy as np

BACKBONE_STRIDES = (4, 8, 16, 32, 64)
ANCHOR_SCALES = (32, 64, 128, 256, 512)
ANCHOR_RATIOS = (0.5, 1.0, 2.0)


def pyramid_shapes(image_h, image_w):
    return [(int(math.ceil(image_h / s)), int(math.ceil(image_w / s))) for s in BACKBONE_STRIDES]


def generate_anchors(image_h=1024, image_w=1024):
    """anchors.bin content: (N,4) f32 normalised (y1,x1,y2,x2), level-major, then y, x, ratio.

    The reference ships this file pre-computed (MaskRCNNConfig.swift:14 TODO) from
    the Matterport-style generator of its un-vendored Keras package (SURVEY.md
    Appendix B): scales 32..512 on strides 4..64, ratios 0.5/1/2, anchor stride 1,
    normalised with (box - [0,0,1,1]) / [h-1, w-1, h-1, w-1].
    """
    out = []
    for (fh, fw), stride, scale in zip(pyramid_shapes(image_h, image_w), BACKBONE_STRIDES, ANCHOR_SCALES):
        ratios = np.asarray(ANCHOR_RATIOS, dtype=np.float64)
        hs = scale / np.sqrt(ratios)
        ws = scale * np.sqrt(ratios)
        cy = (np.arange(fh, dtype=np.float64) * stride)[:, None, None]
        cx = (np.arange(fw, dtype=np.float64) * stride)[None, :, None]
        y1 = np.broadcast_to(cy - 0.5 * hs, (fh, fw, 3))
        x1 = np.broadcast_to(cx - 0.5 * ws, (fh, fw, 3))
        y2 = np.broadcast_to(cy + 0.5 * hs, (fh, fw, 3))
        x2 = np.broadcast_to(cx + 0.5 * ws, (fh, fw, 3))
        out.append(np.stack([y1, x1, y2, x2], axis=-1).reshape(-1, 4))
    boxes = np.concatenate(out, axis=0)
    scale = np.array([image_h - 1, image_w - 1, image_h - 1, image_w - 1], dtype=np.float64)
    shift = np.array([0, 0, 1, 1], dtype=np.float64)
    return ((boxes - shift) / scale).astype(np.float32)


def _iou_matrix(a, b):
    ay1, ax1, ay2, ax2 = [a[:, i:i + 1] for i in range(4)]
    by1, bx1, by2, bx2 = [b[None, :, i] for i in range(4)]
    ih = np.clip(np.minimum(ay2, by2) - np.maximum(ay1, by1), 0, None)
    iw = np.clip(np.minimum(ax2, bx2) - np.maximum(ax1, bx1), 0, None)
    inter = ih * iw
    ua = (ay2 - ay1) * (ax2 - ax1) + (by2 - by1) * (bx2 - bx1) - inter
    return inter / np.maximum(ua, 1e-12)


def rpn_outputs(anchors, image_index=0, image_size=1024, std=(0.1, 0.1, 0.2, 0.2)):
    """(probs (N,2), deltas (N,4)) f32: clustered proposals around K random objects."""
    rng = np.random.default_rng(20260 + image_index)
    n = anchors.shape[0]
    k = int(rng.poisson(12)) + 1
    side = np.exp(rng.uniform(np.log(16.0), np.log(min(700.0, image_size * 0.7)), k))
    aspect = np.exp(rng.uniform(np.log(0.5), np.log(2.0), k))
    h = np.clip(side / np.sqrt(aspect), 4, image_size - 2) / image_size
    w = np.clip(side * np.sqrt(aspect), 4, image_size - 2) / image_size
    cy = rng.uniform(h / 2, 1 - h / 2)
    cx = rng.uniform(w / 2, 1 - w / 2)
    objs = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], axis=1)
    a = anchors.astype(np.float64)
    iou = _iou_matrix(a, objs)
    best = iou.argmax(axis=1)
    iou_max = iou[np.arange(n), best]
    logit = 6.0 * iou_max - 4.0 + rng.normal(0, 0.5, n)
    z = np.stack([-logit / 2, logit / 2], axis=1)
    z -= z.max(axis=1, keepdims=True)
    e = np.exp(z)
    probs = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
    g = objs[best]
    ah, aw = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
    acy, acx = a[:, 0] + 0.5 * ah, a[:, 1] + 0.5 * aw
    gh, gw = g[:, 2] - g[:, 0], g[:, 3] - g[:, 1]
    gcy, gcx = g[:, 0] + 0.5 * gh, g[:, 1] + 0.5 * gw
    enc = np.stack([(gcy - acy) / ah, (gcx - acx) / aw, np.log(gh / ah), np.log(gw / aw)], axis=1) / np.asarray(std)
    matched = iou_max > 0.3
    deltas = np.where(matched[:, None], enc + rng.normal(0, 0.3, (n, 4)), rng.normal(0, 0.5, (n, 4)))
    return probs, deltas.astype(np.float32)


def feature_maps(image_index=0, image_h=1024, image_w=1024, channels=256, dtype=np.float32):
    """P2..P5 as CHW N(0,1) maps (never exactly zero -- matters for quirk Q9)."""
    rng = np.random.default_rng(20260 + image_index)
    maps = []
    for (fh, fw) in pyramid_shapes(image_h, image_w)[:4]:
        m = rng.standard_normal((channels, fh, fw), dtype=np.float32)
        m[m == 0] = 1e-3
        maps.append(m.astype(dtype))
    return maps


def random_rois(n, image_index=0, min_px=8.0, max_px=900.0, image_size=1024, n_pad=0):
    """(n,4) f32 normalised boxes with log-uniform sides; the last n_pad rows are zero (padding)."""
    rng = np.random.default_rng(20260 + image_index)
    side = np.exp(rng.uniform(np.log(min_px), np.log(max_px), n))
    aspect = np.exp(rng.uniform(np.log(0.5), np.log(2.0), n))
    h = np.clip(side / np.sqrt(aspect), 2, image_size - 1) / image_size
    w = np.clip(side * np.sqrt(aspect), 2, image_size - 1) / image_size
    y1 = rng.uniform(0, 1 - h)
    x1 = rng.uniform(0, 1 - w)
    rois = np.stack([y1, x1, y1 + h, x1 + w], axis=1).astype(np.float32)
    rois = np.clip(rois, 0.0, 1.0)
    if n_pad:
        rois[n - n_pad:] = 0.0
    return rois


def classifier_outputs(num_rois, image_index=0, num_classes=81, fg_fraction=0.3):
    """(probabilities (R,ncls), bounding_boxes (R,ncls*4)) f32 with Zipf-distributed foreground classes."""
    rng = np.random.default_rng(20260 + image_index)
    ranks = np.arange(1, num_classes)
    pz = 1.0 / ranks
    pz /= pz.sum()
    fg = rng.uniform(size=num_rois) < fg_fraction
    cls = np.where(fg, rng.choice(ranks, size=num_rois, p=pz), 0)
    score = np.where(fg, rng.uniform(0.7, 1.0, num_rois), rng.uniform(0.4, 0.99, num_rois))
    probs = np.empty((num_rois, num_classes), dtype=np.float64)
    rest = rng.uniform(size=(num_rois, num_classes))
    rest[np.arange(num_rois), cls] = 0
    rest = rest / rest.sum(axis=1, keepdims=True) * (1 - score)[:, None]
    probs[:] = rest
    probs[np.arange(num_rois), cls] = score
    # make the arg-max unambiguous: every other entry strictly below the winner
    probs = np.minimum(probs, (score * 0.999)[:, None])
    probs[np.arange(num_rois), cls] = score
    bbox = rng.standard_normal((num_rois, num_classes * 4))
    return probs.astype(np.float32), bbox.astype(np.float32)
