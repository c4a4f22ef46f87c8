"""The dense graphs in Keras semantics, from UNFOLDED parameters -- a second restatement that cross-checks
oracle/dense_ref.py and weights.fold().

TEST INFRASTRUCTURE ONLY (tests/ import it; the product never does).  PARITY UNPINNED as everything in oracle/: the
reference runs these graphs inside Core ML from .mlmodel files that are not in its tree; they are converted
(Conversion/task.py) from the Matterport Keras model of its un-vendored `maskrcnn` package (requirements.txt:4).  This
file follows that model's graph builders layer by layer (resnet_graph / conv_block / identity_block, the FPN top-down
path, rpn_graph, fpn_classifier_graph, build_fpn_mask_graph) with the Keras layer semantics spelled
out -- ZeroPadding2D + 'valid' convolution, 'same' padding with the extra pixel at the bottom / right, BatchNormalization
as (x - mean) / sqrt(var + 1e-3) * gamma + beta AFTER the biased convolution, Conv2DTranspose with a (kh, kw, out, in)
kernel, Dense on the squeezed 1x1 map -- in float64, from the reference-layout parameter dict
({name: {"kernel" HWIO, "bias", "bn": (gamma, beta, mean, var)}}, what weights.synthetic() produces and the importers
read out of real artefacts).  dense_ref.py evaluates the FOLDED fp16 weights the CUDA pipeline uses; the two must agree
to within the fp16 rounding of the weights (tests/test_dense_ref_cpu.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-3          # keras.layers.BatchNormalization default epsilon


def _t(a):
    return torch.from_numpy(np.asarray(a, np.float64))


class KerasModel:
    def __init__(self, params, architecture=101, num_classes=81):
        self.p, self.arch, self.ncls = params, architecture, num_classes

    # ---- Keras layers ------------------------------------------------------------------------------------------------
    def conv2d(self, x, name, strides=1, padding="valid", out_slice=None):
        """x NCHW; kernel HWIO -> torch OIHW.  'same' = TensorFlow's rule: total padding max((ceil(n/s)-1)*s + k - n, 0),
        the smaller half in front."""
        k, b = _t(self.p[name]["kernel"]), _t(self.p[name]["bias"])
        if out_slice is not None:
            k, b = k[..., out_slice], b[out_slice]
        w = k.permute(3, 2, 0, 1).contiguous()
        if padding == "same":
            pads = []
            for n, ks in ((x.shape[3], k.shape[1]), (x.shape[2], k.shape[0])):       # F.pad order: W first, then H
                total = max((-(-n // strides) - 1) * strides + ks - n, 0)
                pads += [total // 2, total - total // 2]
            x = F.pad(x, pads)
        return F.conv2d(x, w, b, stride=strides)

    def batch_norm(self, x, name):
        gamma, beta, mean, var = (_t(v).view(1, -1, 1, 1) for v in self.p[name]["bn"])
        return (x - mean) / torch.sqrt(var + EPS) * gamma + beta

    def cbr(self, x, name, relu=True, **kw):
        x = self.batch_norm(self.conv2d(x, name, **kw), name)
        return torch.relu(x) if relu else x

    # ---- resnet_graph ------------------------------------------------------------------------------------------------
    def block(self, x, prefix, first, strides):
        y = self.cbr(x, prefix + ".2a", strides=strides if first else 1)
        y = self.cbr(y, prefix + ".2b", padding="same")
        y = self.cbr(y, prefix + ".2c", relu=False)
        shortcut = self.cbr(x, prefix + ".1", relu=False, strides=strides) if first else x     # conv_block vs identity_block
        return torch.relu(y + shortcut)

    def backbone(self, rgb_u8, mean=(123.7, 116.8, 103.9)):
        """rgb [B,H,W,3] u8 -> ([P2..P5] NCHW, rpn probs [B,N,2], rpn deltas [B,N,4]); anchors level-major, then y, x, anchor."""
        x = (_t(rgb_u8.astype(np.float64)) - _t(mean)).permute(0, 3, 1, 2)
        x = F.pad(x, (3, 3, 3, 3))                                                   # ZeroPadding2D((3, 3))
        x = self.cbr(x, "conv1", strides=2)
        x = F.max_pool2d(F.pad(x, (0, 1, 0, 1), value=float("-inf")), 3, 2)          # MaxPooling2D((3,3), strides 2, 'same')
        cs = []
        for s, nb in enumerate({101: (3, 4, 23, 3), 50: (3, 4, 6, 3)}[self.arch]):
            for i in range(nb):
                x = self.block(x, f"res{s + 2}.{i}", first=(i == 0), strides=1 if s == 0 else 2)
            cs.append(x)
        p5 = self.conv2d(cs[3], "fpn.c5p5")
        p4 = F.interpolate(p5, scale_factor=2, mode="nearest") + self.conv2d(cs[2], "fpn.c4p4")   # UpSampling2D + Add
        p3 = F.interpolate(p4, scale_factor=2, mode="nearest") + self.conv2d(cs[1], "fpn.c3p3")
        p2 = F.interpolate(p3, scale_factor=2, mode="nearest") + self.conv2d(cs[0], "fpn.c2p2")
        ps = [self.conv2d(m, f"fpn.p{l}", padding="same") for l, m in zip((2, 3, 4, 5), (p2, p3, p4, p5))]
        p6 = ps[3][:, :, ::2, ::2]                                                   # MaxPooling2D(pool_size=(1,1), strides=2)
        probs, deltas = [], []
        for fmap in ps + [p6]:                                                       # rpn_graph on every level
            shared = torch.relu(self.conv2d(fmap, "rpn.shared", padding="same"))
            logits = self.conv2d(shared, "rpn.head", out_slice=slice(0, 6)).permute(0, 2, 3, 1)      # rpn_class_raw
            bbox = self.conv2d(shared, "rpn.head", out_slice=slice(6, 18)).permute(0, 2, 3, 1)       # rpn_bbox_pred
            b = logits.shape[0]
            probs.append(torch.softmax(logits.reshape(b, -1, 2), dim=-1))
            deltas.append(bbox.reshape(b, -1, 4))
        return ps, torch.cat(probs, 1), torch.cat(deltas, 1)

    # ---- fpn_classifier_graph (after PyramidROIAlign) -----------------------------------------------------------------
    def classifier(self, pooled_nhwc):
        """pooled [M,P,P,256] -> (probabilities [M,ncls], bounding_boxes [M,ncls*4] class-major)."""
        x = _t(pooled_nhwc).permute(0, 3, 1, 2)
        x = self.cbr(x, "cls.conv1")                                                 # Conv2D(1024, (P,P), 'valid') + BN + ReLU
        x = self.cbr(x, "cls.conv2")
        logits = self.conv2d(x, "cls.fc", out_slice=slice(0, self.ncls))[:, :, 0, 0]                 # Dense on the squeezed map
        bbox = self.conv2d(x, "cls.fc", out_slice=slice(self.ncls, 5 * self.ncls))[:, :, 0, 0]
        return torch.softmax(logits, dim=-1), bbox

    # ---- build_fpn_mask_graph (after PyramidROIAlign) ---------------------------------------------------------------------
    def mask(self, pooled_nhwc):
        """pooled [M,P,P,256] -> masks [M,ncls,2P,2P] (sigmoid)."""
        x = _t(pooled_nhwc).permute(0, 3, 1, 2)
        for i in range(1, 5):
            x = self.cbr(x, f"mask.conv{i}", padding="same")
        k, b = _t(self.p["mask.deconv"]["kernel"]), _t(self.p["mask.deconv"]["bias"])       # (kh, kw, out, in)
        x = torch.relu(F.conv_transpose2d(x, k.permute(3, 2, 0, 1).contiguous(), b, stride=2))
        return torch.sigmoid(self.conv2d(x, "mask.final"))
