"""CPU: the multi-GPU exchange (packed detection all_gather) with world_size-2 gloo processes, and shard arithmetic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from centernet_lightning_b200 import distributed as cdist


def test_shard_range_covers_batch():
    for total, world in [(256, 8), (32, 1), (10, 4), (7, 8)]:
        got = []
        for r in range(world):
            s, e = cdist.shard_range(total, r, world)
            got += list(range(s, e))
        assert got == list(range(total))


def test_pack_roundtrip_is_bit_exact():
    g = torch.Generator().manual_seed(0)
    det = {"boxes": torch.randn((3, 5, 4), generator=g), "scores": torch.rand((3, 5), generator=g),
           "labels": torch.randint(0, 65535, (3, 5), generator=g), "embeddings": torch.randn((3, 5, 8), generator=g)}
    back = cdist.unpack_detections(cdist.pack_detections(det))
    for k in det:
        assert torch.equal(back[k], det[k]), k
    assert back["labels"].dtype == torch.int64


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    det = {"boxes": torch.randn((2, 4, 4), generator=g), "scores": torch.rand((2, 4), generator=g),
           "labels": torch.randint(0, 80, (2, 4), generator=g)}
    full = cdist.gather_detections(det)
    q.put((rank, {k: v.numpy() for k, v in full.items()}, {k: v.numpy() for k, v in det.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_detections_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in ("boxes", "scores", "labels"):
        expect = np.concatenate([res[0][2][k], res[1][2][k]], axis=0)       # rank order = batch order
        assert np.array_equal(res[0][1][k], expect) and np.array_equal(res[1][1][k], expect)
