"""Image-level data parallelism for MaskRCNN.predict: images shard across ranks (one process per GPU), every rank
runs the whole pipeline on its shard, and ONE all-gather of the packed (detections | masks) rows returns every
image's result to every rank.  The reference is single-process (EvaluateCommand.swift:166-194 loops over images);
this is the only exchange step of the path.

On GPUs the all-gather is issued by libmaskrcnn_cuda.so itself (mrcnn_predict_allgather: ncclAllGather on the
context's stream, right behind the mask head).  The pure-host helpers below define the sharding and the payload
layout; they are what the world_size-2 gloo tests exercise on CPU.
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import check, lib

DET_ROW = 6


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` images for `rank`; the first total % world ranks get one extra image."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def payload_floats(max_detections=100, mask_size=28):
    """Floats per image in the all-gather payload: D*6 detections followed by D*S*S mask values."""
    return max_detections * DET_ROW + max_detections * mask_size * mask_size


def pack_rows(detections, masks):
    """(B,D,6), (B,D,S,S) -> (B, D*6 + D*S*S): the row layout of mrcnn_predict_allgather's send buffer."""
    b = detections.shape[0]
    return np.concatenate([np.asarray(detections, np.float32).reshape(b, -1), np.asarray(masks, np.float32).reshape(b, -1)], axis=1)


def unpack_rows(packed, max_detections=100, mask_size=28):
    packed = np.asarray(packed, np.float32)
    b = packed.shape[0]
    nd = max_detections * DET_ROW
    return (packed[:, :nd].reshape(b, max_detections, DET_ROW),
            packed[:, nd:].reshape(b, max_detections, mask_size, mask_size))


def all_gather_results(detections, masks, group=None):
    """Host-side equivalent of the library's exchange step over any torch.distributed backend (gloo on CPU):
    every rank contributes the same number of images; returns rank-major concatenated (detections, masks)."""
    import torch
    import torch.distributed as dist
    d, s = detections.shape[1], masks.shape[-1]
    send = torch.from_numpy(pack_rows(detections, masks))
    recv = [torch.empty_like(send) for _ in range(dist.get_world_size(group))]
    dist.all_gather(recv, send, group=group)
    return unpack_rows(torch.cat(recv, 0).numpy(), d, s)


class DistributedMaskRCNN:
    """One rank of a data-parallel MaskRCNN: predict_all() takes this rank's images and returns the results of ALL
    ranks (rank-major), exchanged by a single NCCL all-gather inside the library."""

    def __init__(self, model, rank, world, unique_id=None):
        self.model, self.rank, self.world = model, rank, world
        if world > 1:
            if unique_id is None:
                raise _cabi.MaskRCNNError(_cabi.EINVAL, "unique_id (128 bytes from nccl_unique_id() on rank 0) is required")
            check(model.ctx.handle, lib().mrcnn_comm_init(model.ctx.handle, bytes(unique_id), rank, world))
            model.ctx.nranks = world          # MaskRCNN.submit(allgather=True) sizes its output checks with it

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        check(None, lib().mrcnn_nccl_unique_id(buf))
        return bytes(buf.raw)

    def predict_all(self, images, detections_all=None, masks_all=None):
        m = self.model
        b = images.shape[0]
        if self.world == 1:
            return m.prediction_batch(images, detections_all, masks_all)
        tot = b * self.world
        if detections_all is None:
            detections_all = np.empty((tot, m.D, 6), np.float32)
        if masks_all is None:
            masks_all = np.empty((tot, m.D, m.S, m.S), np.float32)
        check(m.ctx.handle, lib().mrcnn_predict_allgather(m.ctx.handle, b, _cabi.ptr(images), _cabi.ptr(detections_all),
                                                         _cabi.ptr(masks_all)))
        return detections_all, masks_all
