"""Multi-GPU use of the hot path: clips shard embarrassingly, one process per GPU, and the only
exchange step is the final gather of the per-rank (B_local, T, cc) result blocks
(BASELINE config 5).  `torch.distributed` (NCCL on GPUs, gloo in CPU tests) is
the plumbing; no collective runs inside the transform itself.

Results are independent of the world size: the fused kernel's tile schedule and reduction trees do
not depend on the batch, so rank r's block equals rows [lo, hi) of the single-GPU result bit for bit.
"""
from __future__ import annotations


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous block of clips owned by `rank`: sizes differ by at most one, earlier ranks get the extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_blocks(local, total: int, group=None):
    """All-gather variable-size leading-dim blocks into one (total, ...) tensor on every rank.
    Uses all_gather_into_tensor when blocks are equal (one NCCL call), else a padded gather."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {sizes[rank]}")
    if len(set(sizes)) == 1:
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def mfcc_sharded(bft, clips, cc_num=13, rectify_type=0, total=None, group=None, compute=None):
    """Each rank passes ITS shard `clips` (B_local, L); returns the gathered (B_total, T, cc) MFCCs.
    `compute` defaults to `bft.mfcc_batch` (the fused CUDA path); tests inject a stand-in to
    exercise the sharding and gather logic without a GPU."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if total is None:
        total = clips.shape[0] * world
    fn = compute if compute is not None else (lambda x: bft.mfcc_batch(x, cc_num, rectify_type))
    return gather_blocks(fn(clips), total, group)


class OverlappedGather:
    """Chunked compute + all-gather with the collective of chunk k running on a side stream while chunk k+1
    is being transformed (the gather costs the same order as the compute at B200 speed).

    out[k] is the gathered (world * B_chunk, T, cc) block of chunk k: rank r's clips of that chunk sit at rows
    [r * B_chunk, (r + 1) * B_chunk).  Buffers are allocated once and reused, so steady-state steps allocate nothing."""

    def __init__(self, chunks: int = 4, group=None):
        import torch
        self.chunks, self.group = chunks, group
        self.comm = torch.cuda.Stream()
        self.gathered = None
        self.done = [torch.cuda.Event() for _ in range(chunks)]

    def __call__(self, compute, clips):
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(self.group)
        B = clips.shape[0]
        per = (B + self.chunks - 1) // self.chunks
        cur = torch.cuda.current_stream()
        outs = []
        for k in range(self.chunks):
            lo, hi = k * per, min(B, (k + 1) * per)
            if lo >= hi:
                break
            o = compute(clips[lo:hi])
            outs.append(o)
            self.done[k].record(cur)
        if self.gathered is None or len(self.gathered) != len(outs) or self.gathered[0].shape[1:] != outs[0].shape[1:]:
            self.gathered = [torch.empty((world * o.shape[0],) + tuple(o.shape[1:]), dtype=o.dtype, device=o.device) for o in outs]
        with torch.cuda.stream(self.comm):
            for k, o in enumerate(outs):
                self.comm.wait_event(self.done[k])
                dist.all_gather_into_tensor(self.gathered[k], o, group=self.group)
                o.record_stream(self.comm)
        cur.wait_stream(self.comm)
        return self.gathered
