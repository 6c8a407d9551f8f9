"""world_size-2 gloo test of the multi-GPU host logic (sharding bounds + result gather) on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audioflux_b200.dist import gather_blocks, mfcc_sharded, shard_bounds


def test_shard_bounds_partition():
    for total in (0, 1, 7, 8, 1024, 8191):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans[:-1], spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(total * 3 * 2, dtype=torch.float32).reshape(total, 3, 2)
        lo, hi = shard_bounds(total, world, rank)
        got = gather_blocks(full[lo:hi].clone(), total)
        ok1 = torch.equal(got, full)
        # stand-in compute: a deterministic per-clip function, as the real kernel is (batch-size independent)
        clips = torch.arange(total * 5, dtype=torch.float32).reshape(total, 5)
        comp = lambda x: torch.stack([x.sum(1), x.max(1).values], dim=1).unsqueeze(1)   # (B, T=1, cc=2)
        out = mfcc_sharded(None, clips[lo:hi], total=total, compute=comp)
        ok2 = torch.equal(out, comp(clips))
        q.put((rank, bool(ok1), bool(ok2)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_world2_gloo(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert res == [(0, True, True), (1, True, True)]
