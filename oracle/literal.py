"""Statement-level numpy model of the reference's Swift custom layers ("strict" mode of SURVEY.md Appendix A).

TEST INFRASTRUCTURE ONLY (tests/ may import it; nothing under mask-rcnn-coreml_b200/ does).  PARITY UNPINNED: the
reference cannot run here and holds no golden vectors; this file does not change that.  What it adds is a SECOND,
independently written restatement that keeps the reference's own structure -- the Accelerate calls one by one
(cblas_scopy, vDSP_vindex with Float-typed indices, vDSP_vsmul / vsadd / vclip / vthres / vcmprs / maxvi), the
`nonMaxSupression` loop over CGRects, the `Set` of class ids, the ROIAlign item -> group -> batch -> output-item
bookkeeping -- including the behaviour Appendix A calls bugs (Q4 index range, Q5 unflushed last group, padding groups
that clear `count` floats instead of `count` blocks).  oracle.c restates the same semantics in the "intended" mode with
fused loops; tests/test_literal_cpu.py requires the two to agree bit for bit wherever the reference is well defined, and
pins down exactly where (and only where) `intended` departs from the literal behaviour.

Every function cites the Swift lines it models (paths under Sources/Mask-RCNN-CoreML/ of the reference).
Undefined behaviour of the reference is explicit here: `SwiftTrap` for an array index out of range, a `written` mask
for output memory the reference never touches, `class_order` for the iteration order of a Swift `Set`.
"""
import math

import numpy as np

F = np.float32


class SwiftTrap(Exception):
    """The reference would stop the process here (array index out of range, force-unwrap of nil)."""


# ---- Accelerate primitives, as documented by Apple (the library itself is closed) --------------------------------------
def cblas_scopy(n, x, incx, y, incy, x0=0, y0=0):
    """y[y0 + i*incy] = x[x0 + i*incx], i < n."""
    y[y0:y0 + n * incy:incy] = x[x0:x0 + n * incx:incx][:n]


def vDSP_vindex(a, b):
    """C[n] = A[trunc(B[n])] with Float-typed indices B (Utils.swift:28-38)."""
    return a[np.trunc(b).astype(np.int64)].astype(F)


def vDSP_vsorti_descending(keys):
    """vDSP_vsorti(..., order = -1) (Utils.swift:56-66).  Tie order is undocumented (Q1); `intended` = lower index
    first, which a stable sort of the negated keys gives."""
    return np.argsort(-keys.astype(F), kind="stable").astype(np.uint64)


def vDSP_vsmul(a, stride, scalar, n, a0=0):
    a[a0:a0 + n * stride:stride] = (a[a0:a0 + n * stride:stride] * F(scalar)).astype(F)


def vDSP_vsadd(a, stride, scalar, n, a0=0):
    a[a0:a0 + n * stride:stride] = (a[a0:a0 + n * stride:stride] + F(scalar)).astype(F)


def vDSP_vclip(a, lo, hi):
    return np.minimum(np.maximum(a, F(lo)), F(hi)).astype(F)


# ---- Utils.swift ------------------------------------------------------------------------------------------------------------
def strided_slice(p, begin, count, stride, length=1):
    """Utils.swift:17-26."""
    result = np.zeros(count * length, F)
    for l in range(length):
        cblas_scopy(count, p, stride, result, length, x0=begin + l, y0=l)
    return result


def broadcasted_indices(indices, element_length):
    """Utils.swift:112-148: idx -> element_length*idx + {0..element_length-1}, computed in Float."""
    n = indices.size
    result = np.zeros(element_length * n, F)
    for i in range(element_length):
        cblas_scopy(n, indices, 1, result, element_length, y0=i)
    vDSP_vsmul(result, 1, F(element_length), element_length * n)
    for i in range(1, element_length):
        vDSP_vsadd(result, element_length, F(i), n, a0=i)
    return result


def element_wise_multiply(matrix, vector, height, width):
    """Utils.swift:173-181: column i of the (height, width) matrix times vector[i], in place."""
    for i in range(width):
        vDSP_vsmul(matrix, width, vector[i], height, a0=i)


class CGRect:
    """CGRect(anchorDatum:) (Utils.swift:222-231): CGFloat = Double; width/height/min/max of the standardised rect."""

    def __init__(self, datum):
        y1, x1, y2, x2 = (float(v) for v in datum)
        self.x, self.y, self.w, self.h = x1, y1, x2 - x1, y2 - y1

    width = property(lambda s: abs(s.w))
    height = property(lambda s: abs(s.h))
    minX = property(lambda s: s.x if s.w >= 0 else s.x + s.w)
    maxX = property(lambda s: s.x + s.w if s.w >= 0 else s.x)
    minY = property(lambda s: s.y if s.h >= 0 else s.y + s.h)
    maxY = property(lambda s: s.y + s.h if s.h >= 0 else s.y)


def IOU(a, b):
    """Utils.swift:233-246: Double arithmetic, result converted to Float."""
    area_a = a.width * a.height
    if area_a <= 0:
        return F(0)
    area_b = b.width * b.height
    if area_b <= 0:
        return F(0)
    ix0, iy0 = max(a.minX, b.minX), max(a.minY, b.minY)
    ix1, iy1 = min(a.maxX, b.maxX), min(a.maxY, b.maxY)
    inter = max(iy1 - iy0, 0.0) * max(ix1 - ix0, 0.0)
    return F(inter / (area_a + area_b - inter))


def non_max_supression(boxes, indices, iou_threshold, max_):
    """Utils.swift:185-218.  `boxes[index*4 ..< index*4+4]` traps when the slice leaves the array."""
    thr = F(iou_threshold)
    selected = []
    rects = {}

    def rect(i):
        if i not in rects:
            if i * 4 + 4 > boxes.size or i < 0:
                raise SwiftTrap(f"boxes[{i * 4}..<{i * 4 + 4}] out of range (count {boxes.size})")
            rects[i] = CGRect(boxes[i * 4:i * 4 + 4])
        return rects[i]

    for index in indices:
        if len(selected) >= max_:
            return selected
        a = rect(index)
        should_select = a.width > 0 and a.height > 0
        if should_select:
            for j in selected:
                if IOU(a, rect(j)) > thr:
                    should_select = False
                    break
        if should_select:
            selected.append(index)
    return selected


# ---- BoxUtils.swift -----------------------------------------------------------------------------------------------------------
def _expf(x):
    """`exp(Float)`: the arithmetic contract of this repo evaluates it as (float)exp((double)x)."""
    return np.array([math.exp(float(v)) for v in x], np.float64).astype(F)


def apply_box_deltas(boxes, deltas):
    """BoxUtils.swift:34-69, every Float operation rounded on its own, in place on `boxes` (flat, 4 per box)."""
    b = boxes.reshape(-1, 4)
    d = deltas.reshape(-1, 4)
    y1, x1, y2, x2 = b[:, 0].copy(), b[:, 1].copy(), b[:, 2].copy(), b[:, 3].copy()
    half = F(0.5)
    height = (y2 - y1).astype(F)
    width = (x2 - x1).astype(F)
    center_y = (y1 + (half * height).astype(F)).astype(F)
    center_x = (x1 + (half * width).astype(F)).astype(F)
    center_y = (center_y + (d[:, 0] * height).astype(F)).astype(F)
    center_x = (center_x + (d[:, 1] * width).astype(F)).astype(F)
    height = (height * _expf(d[:, 2])).astype(F)
    width = (width * _expf(d[:, 3])).astype(F)
    ry1 = (center_y - (half * height).astype(F)).astype(F)
    rx1 = (center_x - (half * width).astype(F)).astype(F)
    b[:, 0], b[:, 1] = ry1, rx1
    b[:, 2], b[:, 3] = (ry1 + height).astype(F), (rx1 + width).astype(F)


# ---- ProposalLayer.swift:103-195 ---------------------------------------------------------------------------------------------
def proposal_evaluate(probs, deltas, anchors, std=(0.1, 0.1, 0.2, 0.2), pre_nms=6000, max_proposals=1000, iou_thr=0.7,
                      fix_q4=False):
    """One evaluate().  Returns (output (max_proposals, 4), resultIndices, sortedProbabilityIndices).
    fix_q4=False keeps `indices: Array(0 ..< sortedAnchors.count)` (:169-170, 4x too long) and raises SwiftTrap when
    fewer than max_proposals boxes survive among the first n; fix_q4=True iterates 0 ..< n (`intended`)."""
    class_probabilities = np.ascontiguousarray(probs, F).reshape(-1)
    anchor_deltas = np.ascontiguousarray(deltas, F).reshape(-1)
    anchor_data = np.ascontiguousarray(anchors, F).reshape(-1)
    total = class_probabilities.size // 2
    n = min(total, pre_nms)                                                              # :119-120
    object_probabilities = strided_slice(class_probabilities, 1, total, 2)              # :124
    sorted_probability_indices = vDSP_vsorti_descending(object_probabilities)[:n].astype(F)   # :133 (+ toFloat)
    box_indices = broadcasted_indices(sorted_probability_indices, 4)                    # :140
    sorted_deltas = vDSP_vindex(anchor_deltas, box_indices)                             # :144
    sorted_anchors = vDSP_vindex(anchor_data, box_indices)                              # :146-149
    std_dev = np.asarray(std, F)
    element_wise_multiply(sorted_deltas, std_dev, n, std_dev.size)                      # :158
    apply_box_deltas(sorted_anchors, sorted_deltas)                                     # :162
    sorted_anchors = vDSP_vclip(sorted_anchors, 0.0, 1.0)                               # :163
    indices = range(n) if fix_q4 else range(sorted_anchors.size)                        # :169-170
    result_indices = non_max_supression(sorted_anchors, indices, F(iou_thr), max_proposals)
    output = np.full((max_proposals, 4), np.nan, F)       # Core ML does not clear outputs (:188): NaN = "never written"
    for i, r in enumerate(result_indices):                                              # :181-185
        output[i] = sorted_anchors[r * 4:r * 4 + 4]
    output[len(result_indices):] = 0                                                    # :190-192 padTailWithZeros
    return output, result_indices, sorted_probability_indices


# ---- DetectionLayer.swift:107-276 ---------------------------------------------------------------------------------------------
def indices_of_rois_with_high_scores(scores, threshold):
    """DetectionLayer.swift:238-276.  Mutates `scores` like the reference (vDSP_vthres in place)."""
    thr = F(threshold)
    low_count = int((scores < thr).sum())                         # vDSP_vclipc's count of clipped-low elements
    scores[:] = np.where(scores >= thr, scores, F(0))             # vDSP_vthres: C = A >= B ? A : 0
    ramp = np.arange(scores.size, dtype=F)                        # vDSP_vramp
    indices = np.zeros(scores.size - low_count, F)
    gathered = ramp[scores != 0]                                  # vDSP_vcmprs: A where the gate is non-zero
    if gathered.size > indices.size:
        raise SwiftTrap("vDSP_vcmprs writes past `indices`")
    indices[:gathered.size] = gathered
    return indices


def detection_evaluate(rois, classifications, std=(0.1, 0.1, 0.2, 0.2), max_detections=100, score_thr=0.7, iou_thr=0.3,
                       class_order=sorted, final_sort_reverse_ties=False):
    """One evaluate().  Returns (output (max_detections, 6), roi index of every row).
    class_order: the iteration order of `Set(filteredClass)` (:166-170, unspecified in Swift, Q11).
    final_sort_reverse_ties: Swift 4.2 `sorted` is not stable (Q12); True visits equal scores in reverse order."""
    rois_p = np.ascontiguousarray(rois, F).reshape(-1)
    cls_p = np.ascontiguousarray(classifications, F).reshape(-1)
    count = rois_p.size // 4
    bounding_box_deltas = strided_slice(cls_p, 0, count, 6, length=4)                   # :125
    class_ids = strided_slice(cls_p, 4, count, 6)                                       # :127
    scores = strided_slice(cls_p, 5, count, 6)                                          # :128
    filtered_indices = indices_of_rois_with_high_scores(scores, F(score_thr))           # :131-133
    filtered_indices = np.array([i for i in filtered_indices if class_ids[int(i)] > 0], F)     # :136-140
    box_indices = broadcasted_indices(filtered_indices, 4)                              # :144
    filtered_rois = vDSP_vindex(rois_p, box_indices)
    filtered_scores = vDSP_vindex(scores, filtered_indices)
    filtered_class = vDSP_vindex(class_ids, filtered_indices)
    filtered_deltas = vDSP_vindex(bounding_box_deltas, box_indices)
    std_dev = np.asarray(std, F)
    element_wise_multiply(filtered_deltas, std_dev, filtered_class.size, std_dev.size)  # :157-159
    apply_box_deltas(filtered_rois, filtered_deltas)                                    # :162-163
    filtered_rois = vDSP_vclip(filtered_rois, 0.0, 1.0)                                 # :164
    nms_box_ids = []
    for class_id in class_order(set(float(c) for c in filtered_class)):                 # :166-170
        indices_of_class = [o for o, c in enumerate(filtered_class) if c == F(class_id)]
        nms_box_ids += non_max_supression(filtered_rois, indices_of_class, F(iou_thr), max_detections)
    max_elements = min(len(nms_box_ids), max_detections)                                # :189
    zipped = [(b, filtered_scores[b]) for b in nms_box_ids]
    if final_sort_reverse_ties:
        zipped = zipped[::-1]
    zipped = sorted(zipped, key=lambda t: -float(t[1]))                                 # :199-202 (a.1 > b.1)
    result_indices = [b for b, _ in zipped[:max_elements]]
    output = np.full((max_detections, 6), np.nan, F)
    for i, r in enumerate(result_indices):                                              # :217-224
        output[i, :4] = filtered_rois[r * 4:r * 4 + 4]
        output[i, 4] = filtered_class[r]
        output[i, 5] = filtered_scores[r]
    output[len(result_indices):] = 0                                                    # :228-231
    return output, [int(filtered_indices[r]) for r in result_indices]


# ---- PyramidROIAlignLayer.swift ---------------------------------------------------------------------------------------------
def swift_round(x):
    """Swift `round`: to nearest, halfway cases away from zero."""
    return math.copysign(math.floor(abs(x) + 0.5), x)


def rois_to_input_items(rois, selection_factor=224.0, image_w=1024.0, image_h=1024.0):
    """:351-396.  Item = (offset, None) for padding or (offset, (featureMapIndex, (y1, x1, y2, x2)))."""
    ratio = selection_factor / math.sqrt(image_w * image_h)
    items = []
    for i, row in enumerate(rois):
        y1, x1, y2, x2 = (float(v) for v in row[:4])
        width, height = x2 - x1, y2 - y1
        prod = width * height
        if prod < 0 or math.isnan(prod):
            level_f = float("nan")                      # sqrt of a negative number
        elif prod == 0:
            level_f = float("-inf")                     # log2(0)
        else:
            level_f = math.log2(math.sqrt(prod) / ratio) + 4.0
        valid = not math.isnan(level_f) and not math.isinf(level_f)
        level = 2 if not valid else min(5, max(2, int(swift_round(level_f))))
        items.append((i, (level - 2, tuple(row[:4])) if valid else None))
    return items


def group_input_items_by_content(items, max_compute_batch_size=64):
    """:398-468.  Group = (offset, "padding", count) or (offset, mapIndex, [regions]).  As in the reference there is NO
    close after the loop (Q5): the run that is open when the items end is dropped."""
    results = []
    state = {"offset": 0, "padding": None, "map": None, "regions": None}

    def close_regions():
        if state["map"] is not None and state["regions"] is not None:
            results.append((state["offset"], state["map"], state["regions"]))
            state["offset"] += len(state["regions"])
            state["map"], state["regions"] = None, None

    def close_padding():
        if state["padding"] is not None:
            results.append((state["offset"], "padding", state["padding"]))
            state["offset"] += state["padding"]
            state["padding"] = None

    for _, content in items:
        if content is None:
            close_regions()
            state["padding"] = 1 if state["padding"] is None else state["padding"] + 1
        else:
            map_index, region = content
            close_padding()
            if state["map"] is not None and state["map"] == map_index:
                state["regions"].append(region)
            else:
                close_regions()
                state["map"], state["regions"] = map_index, [region]
            if len(state["regions"]) == max_compute_batch_size:
                close_regions()
    return results


def batch_input_groups(groups, max_compute_batch_size=64):
    """:470-498."""
    def compute_size(g):
        return 0 if g[1] == "padding" else len(g[2])
    batches, in_batch, current = [], [], 0
    for g in groups:
        nxt = current + compute_size(g)
        if nxt <= max_compute_batch_size:
            in_batch.append(g)
            current = nxt
        else:
            batches.append(in_batch)
            in_batch, current = [g], compute_size(g)
    if in_batch:
        batches.append(in_batch)
    return batches


def crop_and_resize_bilinear(fmap, box, pool):
    """MPSNNCropAndResizeBilinear (:212-223; closed source) restated as TensorFlow crop_and_resize, bilinear,
    extrapolation value 0, fp32.  fmap (C,H,W), box (y1,x1,y2,x2) normalised -> (C,pool,pool)."""
    c, h, w = fmap.shape
    y1, x1, y2, x2 = (F(v) for v in box)
    hm1, wm1 = F(h - 1), F(w - 1)
    steps = np.arange(pool, dtype=F)
    if pool > 1:
        hs = F(F((y2 - y1)) * hm1) / F(pool - 1)
        ws = F(F((x2 - x1)) * wm1) / F(pool - 1)
        in_y = (F(y1 * hm1) + (steps * F(hs)).astype(F)).astype(F)
        in_x = (F(x1 * wm1) + (steps * F(ws)).astype(F)).astype(F)
    else:
        in_y = np.full(1, F(F(half_sum(y1, y2)) * hm1), F)
        in_x = np.full(1, F(F(half_sum(x1, x2)) * wm1), F)
    ok_y = ~((in_y < 0) | (in_y > hm1))
    ok_x = ~((in_x < 0) | (in_x > wm1))
    sy, sx = np.where(ok_y, in_y, F(0)), np.where(ok_x, in_x, F(0))
    t, b = np.floor(sy).astype(np.int64), np.ceil(sy).astype(np.int64)
    l, r = np.floor(sx).astype(np.int64), np.ceil(sx).astype(np.int64)
    ly = (sy - np.floor(sy)).astype(F)[None, :, None]
    lx = (sx - np.floor(sx)).astype(F)[None, None, :]
    tl, tr = fmap[:, t][:, :, l], fmap[:, t][:, :, r]
    bl, br = fmap[:, b][:, :, l], fmap[:, b][:, :, r]
    top = (tl + ((tr - tl).astype(F) * lx).astype(F)).astype(F)
    bot = (bl + ((br - bl).astype(F) * lx).astype(F)).astype(F)
    out = (top + ((bot - top).astype(F) * ly).astype(F)).astype(F)
    out[:, ~ok_y, :] = 0
    out[:, :, ~ok_x] = 0
    return out


def half_sum(a, b):
    return F(0.5) * F(a + b)


def pyramid_roialign_evaluate(rois, fmaps, pool, image_w=1024.0, image_h=1024.0, max_batch_size=64, stale=np.nan):
    """One evaluate() (:79-181) with the output-item bookkeeping of performBatch (:186-240) and copyOutput (:245-274).
    Returns (out (R,C,pool,pool), written (same shape, bool)).  `out` starts as `stale` everywhere: whatever stays
    `stale` / unwritten is memory the reference leaves as Core ML handed it over (:103-106 early return, Q5, and
    padding groups clearing `count` floats rather than `count` blocks)."""
    rois = np.ascontiguousarray(rois, F)
    fmaps = [np.ascontiguousarray(f, F) for f in fmaps]
    channels = fmaps[0].shape[0]
    total = rois.shape[0]
    result_stride = channels * pool * pool                                      # outputs[0].strides[0]
    out = np.full(total * result_stride, stale, F)
    written = np.zeros(total * result_stride, bool)
    items = rois_to_input_items(rois, 224.0, image_w, image_h)
    groups = group_input_items_by_content(items, max_batch_size)
    batches = batch_input_groups(groups, max_batch_size)
    for batch in batches:                                                       # empty -> nothing is written (:103-106)
        for offset, kind, payload in batch:
            if kind == "padding":                                               # ROIAlignOutputItem(.padding(count:))
                start = offset * result_stride
                out[start:start + payload] = 0                                  # copyMemory(byteCount: 4*paddingCount)
                written[start:start + payload] = True
            else:
                for r, region in enumerate(payload):                            # one output item per region (:204-226)
                    start = (offset + r) * result_stride
                    out[start:start + result_stride] = crop_and_resize_bilinear(fmaps[kind], region, pool).reshape(-1)
                    written[start:start + result_stride] = True
    shape = (total, channels, pool, pool)
    return out.reshape(shape), written.reshape(shape)


# ---- TimeDistributedClassifierLayer.swift ---------------------------------------------------------------------------------------
def maximum_value_with_index(values):
    """:177-192, vDSP_maxvi: the FIRST maximum."""
    i = int(np.argmax(values))
    return values[i], i


def index_mapping(first_input, remove_zeros):
    """MultiArrayBatchProvider.init (:108-133): with removeZeros only blocks whose EVERY element is non-zero stay (Q9)."""
    mapping = []
    for i in range(first_input.shape[0]):
        if not remove_zeros or bool(np.all(first_input[i] != 0)):
            mapping.append(i)
    return mapping


def classifier_process_output(probabilities, bounding_boxes):
    """:50-88 on the Classifier model's outputs (Double in the reference, converted to Float first, :65-71)."""
    p64 = np.asarray(probabilities, np.float64)
    b64 = np.asarray(bounding_boxes, np.float64)
    out = np.full((p64.shape[0], 6), np.nan, F)
    for actual_index in index_mapping(p64, remove_zeros=False):
        float_buffer = p64[actual_index].astype(F)
        value, last_index = maximum_value_with_index(float_buffer)
        out[actual_index, 4] = F(last_index)
        out[actual_index, 5] = value
        float_buffer = b64[actual_index].astype(F)
        for z in range(4):
            out[actual_index, z] = float_buffer[last_index * 4 + z]
    return out


# ---- TimeDistributedMaskLayer.swift:39-91 ------------------------------------------------------------------------------------------
def mask_evaluate(pooled, detections, masks_of):
    """pooled (D,C,P,P); detections (D,6); masks_of(i) -> the Mask model's output (ncls,S,S) for pooled[i].
    Returns out (D,S,S).  The class is read at the COMPACTED index (detections[stride*i+4], :71, Q14) and the tail is
    zeroed from `resultCount` on (:87-89)."""
    pooled = np.asarray(pooled, F)
    detections = np.asarray(detections, F).reshape(-1, 6)
    mapping = index_mapping(pooled, remove_zeros=True)                          # :52
    out = None
    for i, actual_index in enumerate(mapping):
        masks = np.asarray(masks_of(actual_index), np.float64)
        if out is None:
            out = np.full((pooled.shape[0],) + masks.shape[1:], np.nan, F)
        class_id = int(detections[i, 4])                                        # :71
        out[actual_index] = masks[class_id].astype(F)                           # :73-83
    if out is None:
        s = 2 * pooled.shape[-1]
        out = np.full((pooled.shape[0], s, s), np.nan, F)
    out[len(mapping):] = 0                                                      # :87-89
    return out


# ---- Detection.swift -------------------------------------------------------------------------------------------------------------
def detections_from_feature_value(raw_detections, raw_masks=None):
    """Detection.swift:23-62 and :64-99.  Returns [(index, (x, y, w, h), classId, score, mask_u8 | None)]."""
    det = np.asarray(raw_detections, F).reshape(-1, 6)
    results = []
    for i in range(det.shape[0]):
        score = float(det[i, 5])
        if score > 0.7:                                                         # Double literal (:38)
            class_id = int(det[i, 4])
            y1, x1, y2, x2 = (float(v) for v in det[i, :4])
            mask = None
            if raw_masks is not None and raw_masks.shape[0] > i:
                m = np.asarray(raw_masks[i], np.float64).reshape(-1)
                mask = (255 - (m / 2 * 255)).astype(np.uint8)                   # UInt8(...) truncates (:83-85)
            results.append((i, (x1, y1, x2 - x1, y2 - y1), class_id, score, mask))
    return results
