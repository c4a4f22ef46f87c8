"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding and the all-gather payload layout that
mrcnn_predict_allgather uses on the device (one packed row per image = D*6 detections | D*S*S masks)."""
import os

import numpy as np
import pytest


def _fake_results(image_ids, d=100, s=28):
    det = np.zeros((len(image_ids), d, 6), np.float32)
    masks = np.zeros((len(image_ids), d, s, s), np.float32)
    for k, i in enumerate(image_ids):
        rng = np.random.default_rng(1000 + i)
        n = int(rng.integers(0, d))
        det[k, :n] = rng.uniform(size=(n, 6)).astype(np.float32)
        det[k, :n, 4] = rng.integers(1, 81, n)
        masks[k, :n] = rng.uniform(size=(n, s, s)).astype(np.float32)
    return det, masks


def _worker(rank, world, port, total, out):
    import torch.distributed as dist
    import maskrcnn_b200 as m
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = m.distributed.shard_range(total, rank, world)
    det, masks = _fake_results(list(range(lo, hi)))
    all_det, all_masks = m.distributed.all_gather_results(det, masks)
    want_det, want_masks = _fake_results(list(range(total)))
    ok = np.array_equal(all_det, want_det) and np.array_equal(all_masks, want_masks)
    out.put((rank, bool(ok), all_det.shape, all_masks.shape))
    dist.destroy_process_group()


def test_allgather_payload_two_ranks_gloo(pkg):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, total = 2, 6
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, sd, sm in res:
        assert ok, f"rank {rank}: gathered results differ from the rank-major concatenation"
        assert sd == (total, 100, 6) and sm == (total, 100, 28, 28)


def test_shard_ranges_cover_and_balance(pkg):
    sr = pkg.distributed.shard_range
    for total in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [sr(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sr(64, 3, 8) == (24, 32)          # BASELINE configs[3]: batch 64 over 8 GPUs = 8 images per rank


def test_pack_unpack_roundtrip(pkg):
    det, masks = _fake_results([3, 4, 5])
    packed = pkg.distributed.pack_rows(det, masks)
    assert packed.shape == (3, pkg.distributed.payload_floats()) and packed.shape[1] * 4 == 316000
    d2, m2 = pkg.distributed.unpack_rows(packed)
    np.testing.assert_array_equal(d2, det)
    np.testing.assert_array_equal(m2, masks)
