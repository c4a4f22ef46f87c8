"""Matterport Mask R-CNN Keras weights (`mask_rcnn_coco.h5`, the file Conversion/task.py:163 loads) -> this library's
reference-layout parameters (weights.fold() input) and back.  SURVEY.md section 8 f4.

File layout written by Keras `save_weights`: one group per layer, inside it one dataset per weight under the weight's
own name, e.g. `conv1/conv1/kernel:0`, `bn_conv1/bn_conv1/moving_variance:0`; layers of the nested RPN model sit under
`rpn_model/` (`rpn_model/rpn_conv_shared/kernel:0`).  Datasets are found by their `<layer>/<weight>` suffix, so the
`layer_names` / `weight_names` attributes are not needed.  Kernels: Conv2D HWIO, Dense (in, out), Conv2DTranspose
(kh, kw, out, in) -- the layouts weights.synthetic() uses, so nothing is transposed here.

Read through h5lite (no h5py / libhdf5 in this image); NEVER RUN ON THE REAL FILE -- it is not reachable offline.
"""
import numpy as np

from . import h5lite
from . import weights as W
from .mlmodel import keras_names

_BN = ("gamma:0", "beta:0", "moving_mean:0", "moving_variance:0")
_SPLIT = {"rpn.head": (6, 12)}          # rpn_class_raw | rpn_bbox_pred (3 anchors x 2, 3 anchors x 4)


def _splits(name, num_classes):
    if name == "cls.fc":
        return (num_classes, 4 * num_classes)          # mrcnn_class_logits | mrcnn_bbox_fc
    return _SPLIT.get(name)


def params_from_keras_h5(path_or_bytes, architecture=101, num_classes=81, pool_classifier=7):
    """-> {name: {"kernel", "bias", ["bn", "bn_eps"]}} for weights.fold()."""
    ds = h5lite.read(path_or_bytes).datasets()

    def find(layer, weight):
        suffix = f"{layer}/{weight}"
        hits = [k for k in ds if k == suffix or k.endswith("/" + suffix)]
        if not hits:
            raise KeyError(f"weight '{suffix}' not found in the Keras file ({len(ds)} datasets)")
        if len({ds[h].tobytes() for h in hits}) > 1:
            raise ValueError(f"weight '{suffix}' is ambiguous: {hits}")
        return np.asarray(ds[hits[0]], np.float32)

    names = keras_names(architecture)
    params = {}
    for name, which, kind, kh, kw, cin, cout, has_bn, relu in W.layer_table(architecture, num_classes, pool_classifier):
        convs, bn = names[name]
        ks, bs = [], []
        for c in convs:
            k = find(c, "kernel:0")
            if k.ndim == 2:                             # Dense (in, out) -> 1x1 convolution
                k = k.reshape(1, 1, *k.shape)
            ks.append(k)
            bs.append(find(c, "bias:0"))
        kernel = np.concatenate(ks, axis=2 if kind == "deconv" else -1)
        want = (kh, kw, cout, cin) if kind == "deconv" else (kh, kw, cin, cout)
        if kernel.shape != want:
            raise ValueError(f"{name}: kernel shape {kernel.shape}, expected {want}")
        p = {"kernel": np.ascontiguousarray(kernel), "bias": np.concatenate(bs)}
        if bn is not None:
            p["bn"] = tuple(find(bn, w) for w in _BN)
            p["bn_eps"] = W.BN_EPS                      # Keras BatchNormalization default; not stored in the file
        params[name] = p
    return params


def keras_tree_from_params(params, architecture=101, num_classes=81, pool_classifier=7):
    """The inverse: nested {layer: {layer: {weight: array}}} in the Keras save_weights layout (for h5lite.write)."""
    names = keras_names(architecture)
    tree = {}

    def put(layer, weights):
        top = "rpn_model" if layer.startswith("rpn_") else layer
        tree.setdefault(top, {}).setdefault(layer, {}).update(weights)

    for name, which, kind, kh, kw, cin, cout, has_bn, relu in W.layer_table(architecture, num_classes, pool_classifier):
        convs, bn = names[name]
        p = params[name]
        kernel, bias = np.asarray(p["kernel"], np.float32), np.asarray(p["bias"], np.float32)
        sizes = _splits(name, num_classes) or (cout,)
        start = 0
        for c, n in zip(convs, sizes):
            k = kernel[:, :, start:start + n, :] if kind == "deconv" else kernel[..., start:start + n]
            if c in ("mrcnn_class_logits", "mrcnn_bbox_fc"):
                k = k.reshape(k.shape[2], k.shape[3])   # Dense
            put(c, {"kernel:0": np.ascontiguousarray(k), "bias:0": bias[start:start + n]})
            start += n
        if bn is not None and "bn" in p:
            put(bn, dict(zip(_BN, (np.asarray(x, np.float32) for x in p["bn"]))))
    return tree


def import_products(path_or_bytes, architecture=101, num_classes=81, pool_classifier=7):
    """Keras weight file -> [MaskRCNN, Classifier, Mask] weight blobs (weights.pack_blob)."""
    params = params_from_keras_h5(path_or_bytes, architecture, num_classes, pool_classifier)
    folded = W.fold(params, architecture, num_classes, pool_classifier)
    return [W.pack_blob(which, W.device_tensors(folded, which, architecture, num_classes, pool_classifier))
            for which in (W.MAIN, W.CLASSIFIER, W.MASK)]
