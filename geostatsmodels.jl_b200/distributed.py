"""Multi-GPU plumbing: one process per GPU, targets partitioned, samples replicated.

The reference parallelises over targets with contiguous index chunks and private state per chunk
(src/models.jl:236-247); there is no exchange step on the data path.  Here every rank owns one GPU and
one contiguous slab of the linear target index; the only collective is the one-off broadcast of the
packed samples (NCCL over NVLink on GPUs, gloo in the CPU tests).  Predictions stay on the rank that
produced them (each rank writes its own slice of the output)."""
from __future__ import annotations

import numpy as np


def shard(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous chunk `rank` of `world` over range(total), sizes differing by at most one --
    the policy of ChunkSplitters.index_chunks(inds, n=nthreads) used at models.jl:238."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(int(total), int(world))
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return first, count


def replicate_samples(coords, values, src: int = 0, device=None):
    """Broadcast (coords[n, dim], values[n]) from rank `src` to every rank as ONE packed float64 tensor.
    Returns (coords, values) torch tensors on `device` (the packed buffer is n x (dim + 1))."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = torch.device("cpu") if device is None else torch.device(device)
    shape = torch.zeros(2, dtype=torch.int64, device=dev)
    if rank == src:
        c = np.ascontiguousarray(coords, dtype=np.float64)
        v = np.ascontiguousarray(values, dtype=np.float64)
        shape[0], shape[1] = c.shape[0], c.shape[1]
    if world > 1:
        dist.broadcast(shape, src=src)
    n, dim = int(shape[0]), int(shape[1])
    packed = torch.empty((n, dim + 1), dtype=torch.float64, device=dev)
    if rank == src:
        packed.copy_(torch.from_numpy(np.concatenate([c, v[:, None]], axis=1)))
    if world > 1:
        dist.broadcast(packed, src=src)
    return packed[:, :dim].contiguous(), packed[:, dim].contiguous()
