"""Plain PyTorch fp32 restatement of the dense graphs (ResNet+FPN+RPN, classifier head, mask head) that the
reference runs inside Core ML from its .mlmodel artefacts (not in the reference tree; architecture per SURVEY.md
Appendix B).  Built from the same folded fp16 weights as the CUDA pipeline.  act_half=True rounds activations to
fp16 at the points where the pipeline stores fp16 tensors, so the remaining differences are accumulation order only.

TEST INFRASTRUCTURE ONLY (see oracle.c header): used by tests/ as the checker of the dense stages and by
bench.py's cpu_baseline / --impl reference legs (device="cpu").  PARITY UNPINNED: no artefact or golden output of
the reference's dense graph is available offline."""
import numpy as np
import torch
import torch.nn.functional as F


def _dev(device=None):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    if device is not None:
        return torch.device(device)
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


class Ref:
    def __init__(self, folded, architecture=101, num_classes=81, act_half=True, device=None):
        self.f, self.arch, self.ncls, self.half, self.dev = folded, architecture, num_classes, act_half, _dev(device)
        self._cache = {}

    def r(self, t):
        return t.half().float() if self.half else t

    def conv(self, x, name, stride=1, pad=0, relu=False, res=None, rnd=True):
        if name not in self._cache:
            w, b = self.f[name]
            self._cache[name] = (torch.from_numpy(w.astype(np.float32)).permute(0, 3, 1, 2).contiguous().to(self.dev),
                                 torch.from_numpy(b).to(self.dev))
        wt, bt = self._cache[name]
        y = F.conv2d(x, wt, bt, stride=stride, padding=pad)
        if res is not None:
            y = y + res
        if relu:
            y = torch.relu(y)
        return self.r(y) if rnd else y

    def backbone(self, rgb_u8, mean=(123.7, 116.8, 103.9)):
        """rgb [B,H,W,3] u8 -> ([P2..P5] NHWC fp32 tensors, probs [B,N,2], deltas [B,N,4])."""
        x = torch.from_numpy(rgb_u8.astype(np.float32)).to(self.dev) - torch.tensor(mean, device=self.dev)
        x = self.r(x).permute(0, 3, 1, 2).contiguous()
        x = self.conv(x, "conv1", stride=2, pad=3, relu=True)
        x = F.max_pool2d(F.pad(x, (0, 1, 0, 1), value=float("-inf")), 3, 2)
        nb = {101: (3, 4, 23, 3), 50: (3, 4, 6, 3)}[self.arch]
        cs = []
        for s, n in enumerate(nb):
            for i in range(n):
                stride = 2 if (i == 0 and s > 0) else 1
                p = f"res{s + 2}.{i}"
                t = self.conv(x, p + ".2a", stride=stride, relu=True)
                t = self.conv(t, p + ".2b", pad=1, relu=True)
                res = self.conv(x, p + ".1", stride=stride) if i == 0 else x
                x = self.conv(t, p + ".2c", relu=True, res=res)
            cs.append(x)
        m = [None] * 4
        m[3] = self.conv(cs[3], "fpn.c5p5")
        for l in (2, 1, 0):
            up = F.interpolate(m[l + 1], scale_factor=2, mode="nearest")
            m[l] = self.conv(cs[l], f"fpn.c{l + 2}p{l + 2}", res=up)
        p = [self.conv(m[l], f"fpn.p{l + 2}", pad=1) for l in range(4)]
        p.append(p[3][:, :, ::2, ::2])
        probs, deltas = [], []
        for x in p:
            sh = self.conv(x, "rpn.shared", pad=1, relu=True)
            hd = self.conv(sh, "rpn.head", rnd=False).permute(0, 2, 3, 1)      # [B,h,w,18]
            b = hd.shape[0]
            probs.append(torch.softmax(hd[..., :6].reshape(b, -1, 2), dim=-1))
            deltas.append(hd[..., 6:].reshape(b, -1, 4))
        return [t.permute(0, 2, 3, 1).contiguous() for t in p[:4]], torch.cat(probs, 1), torch.cat(deltas, 1)

    def classifier(self, pooled_nhwc):
        """pooled [M,P,P,256] (fp16-representable) -> (probs [M,ncls], bbox [M,ncls*4], logits)."""
        x = torch.as_tensor(pooled_nhwc, dtype=torch.float32, device=self.dev)
        m = x.shape[0]
        w1, b1 = self.f["cls.conv1"]
        f1 = self.r(torch.relu(x.reshape(m, -1) @ torch.from_numpy(w1.astype(np.float32)).reshape(w1.shape[0], -1).T.to(self.dev)
                               + torch.from_numpy(b1).to(self.dev)))
        w2, b2 = self.f["cls.conv2"]
        f2 = self.r(torch.relu(f1 @ torch.from_numpy(w2.astype(np.float32)).reshape(w2.shape[0], -1).T.to(self.dev)
                               + torch.from_numpy(b2).to(self.dev)))
        w3, b3 = self.f["cls.fc"]
        lg = f2 @ torch.from_numpy(w3.astype(np.float32)).reshape(w3.shape[0], -1).T.to(self.dev) + torch.from_numpy(b3).to(self.dev)
        return torch.softmax(lg[:, :self.ncls], dim=-1), lg[:, self.ncls:], lg

    def mask(self, pooled_nhwc):
        """pooled [M,P,P,256] -> sigmoid masks [M,ncls,2P,2P]."""
        x = torch.as_tensor(pooled_nhwc, dtype=torch.float32, device=self.dev).permute(0, 3, 1, 2).contiguous()
        for i in range(1, 5):
            x = self.conv(x, f"mask.conv{i}", pad=1, relu=True)
        w, b = self.f["mask.deconv"]
        co = w.shape[0] // 4
        wt = torch.from_numpy(w.astype(np.float32)).reshape(2, 2, co, -1).permute(3, 2, 0, 1).contiguous().to(self.dev)
        x = self.r(torch.relu(F.conv_transpose2d(x, wt, torch.from_numpy(b[:co]).to(self.dev), stride=2)))
        return torch.sigmoid(self.conv(x, "mask.final", rnd=False))
