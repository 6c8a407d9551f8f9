"""Worker of tests/test_gpu_multi.py: one rank per GPU under torchrun (NCCL).  Each rank transforms its own seeded
shard with the fused MFCC + all-gather kernel (audioflux_b200.dist.PeerScatter) and checks that EVERY slot of the
gathered array equals the locally recomputed transform of that rank's shard, bit for bit; then the same through the
NCCL path (OverlappedGather)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)

import audioflux_b200 as af  # noqa: E402
from audioflux_b200.dist import PeerScatter, OverlappedGather  # noqa: E402


def shard(rank, B, L, dev):
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    return 0.1 * torch.randn((B, L), generator=g, device=dev, dtype=torch.float32)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    B, L, CC = 24, 48000, 40
    S, D = af.SpectralFilterBankScaleType, af.SpectralDataType
    bft = af.BFT(128, 11, 48000, slide_length=512, scale_type=S.MEL, data_type=D.POWER)
    x = shard(rank, B, L, dev)
    sc = PeerScatter(bft, B, L, CC)
    for rep in range(3):                                   # repeated steps reuse the mapped buffers (alternating slots)
        for arr in sc._gathered:
            arr.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        out = sc(x)
        sc.fence()
        torch.cuda.synchronize()
        for r in range(world):
            want = bft.mfcc_batch(shard(r, B, L, dev), CC)
            assert torch.equal(out[r], want), f"rank {rank}: slot {r} differs (rep {rep})"
    # write-after-read: consecutive steps with NO host synchronisation in between.  Step k's gathered array
    # must still hold step k's blocks after step k+1 has been launched and fenced (double-buffered slots), and the
    # third step reuses the first slot.
    x2 = shard(rank + 50, B, L, dev)
    a = sc(x)
    sc.fence()
    a_copy_later = torch.empty_like(a)
    b = sc(x2)                                             # peers may already be storing step k+1 ...
    a_copy_later.copy_(a)                                  # ... while this rank still reads step k (same stream, before the next fence)
    sc.fence()
    c = sc(x)
    sc.fence()
    torch.cuda.synchronize()
    assert a.data_ptr() != b.data_ptr() and c.data_ptr() == a.data_ptr()
    for r in range(world):
        assert torch.equal(a_copy_later[r], bft.mfcc_batch(shard(r, B, L, dev), CC)), f"rank {rank}: step k block {r} was overwritten"
        assert torch.equal(b[r], bft.mfcc_batch(shard(r + 50, B, L, dev), CC)), f"rank {rank}: step k+1 block {r} differs"
        assert torch.equal(c[r], bft.mfcc_batch(shard(r, B, L, dev), CC)), f"rank {rank}: step k+2 block {r} differs"
    dist.barrier()
    # chunked launches (a shard streamed in from the host): same result
    for arr in sc._gathered:
        arr.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    for lo in range(0, B, 7):
        sc(x[lo:lo + 7].contiguous(), clip_offset=lo)
    sc.fence()
    torch.cuda.synchronize()
    for r in range(world):
        assert torch.equal(sc.gathered[r], bft.mfcc_batch(shard(r, B, L, dev), CC)), f"rank {rank}: chunked slot {r} differs"
    torch.cuda.synchronize()
    dist.barrier()
    sc.close()
    ov = OverlappedGather(chunks=3)
    res = ov(lambda c: bft.mfcc_batch(c, CC), x)
    torch.cuda.synchronize()
    per = B // 3
    for k, o in enumerate(res):
        for r in range(world):
            want = bft.mfcc_batch(shard(r, B, L, dev)[k * per:(k + 1) * per], CC)
            assert torch.equal(o[r * per:(r + 1) * per], want), f"rank {rank}: nccl chunk {k} slot {r} differs"
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("PEER_SCATTER_OK", world)


if __name__ == "__main__":
    main()
