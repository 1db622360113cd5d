"""Multi-GPU plumbing, one PROCESS per GPU (torchrun): targets partitioned, samples replicated.

The reference parallelises over targets with contiguous index chunks and private state per chunk
(src/models.jl:236-247); there is no exchange step on the data path.  Here every rank owns one GPU and
one contiguous slab of the linear target index; the only collective is the broadcast of the samples
(NCCL over NVLink on GPUs, gloo in the CPU tests).  Predictions stay on the rank that produced them (each
rank writes its own slice of the output).  The single-process form of the same scheme -- one caller, all GPUs,
`ncclCommInitAll` inside the C ABI -- is `gsm_multi_*` (csrc/multi.cu, `fitpredict(..., devices=[...])`)."""
from __future__ import annotations

import numpy as np


def shard(total: int, rank: int, world: int, unit: int = 1) -> tuple[int, int]:
    """Contiguous chunk `rank` of `world` over range(total), sizes differing by at most one `unit` --
    the policy of ChunkSplitters.index_chunks(inds, n=nthreads) used at models.jl:238.  `unit` > 1 keeps the
    boundaries on multiples of it (pairs of grid rows / planes: no slab cuts a patch of the solve kernels);
    the remainder goes to the last chunk."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    unit = max(1, int(unit))
    units = int(total) // unit
    base, rem = divmod(units, int(world))
    first = (rank * base + min(rank, rem)) * unit
    count = (base + (1 if rank < rem else 0)) * unit
    if rank == world - 1:
        count = int(total) - first
    return first, count


def grid_unit(dims) -> int:
    """targets in a pair of grid rows (2-D) / planes (3-D)"""
    dims = tuple(int(d) for d in dims)
    if len(dims) < 2:
        return 1
    return 2 * dims[0] * (dims[1] if len(dims) > 2 else 1)


_WARM = set()


def _warm(dev):
    """The first collective on a fresh communicator sets up its channels (tens of ms): do it once, outside any
    timed region, like gsm_multi_create does for the single-process form."""
    import torch
    import torch.distributed as dist
    key = (str(dev), dist.get_world_size())
    if key not in _WARM:
        dist.broadcast(torch.zeros(1 << 18, dtype=torch.float64, device=dev), src=0)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        _WARM.add(key)


def replicate_samples(coords, values, src: int = 0, device=None):
    """Broadcast (coords[n, dim], values[n]) from rank `src` to every rank.  Returns device tensors.

    Rank `src` passes numpy arrays (page-locked ones -- gsm_host_alloc / PinnedArray / torch pinned -- are DMA'd
    straight to the device; pageable ones are staged by the driver); the other ranks pass None.  Two broadcasts of
    the raw buffers: no host-side packing pass over the samples."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = torch.device("cpu") if device is None else torch.device(device)
    shape = torch.zeros(2, dtype=torch.int64, device=dev)
    if rank == src:
        c = np.ascontiguousarray(coords, dtype=np.float64)
        v = np.ascontiguousarray(values, dtype=np.float64)
        if c.ndim == 1:
            c = c[:, None]
        shape[0], shape[1] = c.shape[0], c.shape[1]
    if world > 1:
        _warm(dev)
        dist.broadcast(shape, src=src)
    n, dim = int(shape[0]), int(shape[1])
    coords_d = torch.empty((n, dim), dtype=torch.float64, device=dev)
    values_d = torch.empty((n,), dtype=torch.float64, device=dev)
    if rank == src:
        coords_d.copy_(torch.from_numpy(c), non_blocking=True)
        values_d.copy_(torch.from_numpy(v), non_blocking=True)
    if world > 1:
        dist.broadcast(coords_d, src=src)
        dist.broadcast(values_d, src=src)
    return coords_d, values_d


def fitpredict_sharded(model, data, dom, *, src: int = 0, ctx=None, out=None, name: str = "z", **kw):
    """`fitpredict` over torch.distributed: every rank calls it with the same model / domain / options; the geotable
    lives on rank `src` (the others pass None).  The samples are replicated (replicate_samples), every rank predicts
    its contiguous slab of the target index (aligned to pairs of grid rows / planes) into its own arrays.
    Returns (prediction geotable of the slab, (first, count))."""
    import torch
    import torch.distributed as dist
    from . import api

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if ctx is None:
        ctx = api.default_context()
    dev = torch.device("cuda", ctx.device)
    X = z = None
    if rank == src:
        names = data.names()
        name = names[0]
        z = np.asarray(data.table[name], dtype=np.float64)
        X = api._pointset_of(data.domain)
    coords_d, values_d = replicate_samples(X, z, src=src, device=dev)
    torch.cuda.current_stream(dev).synchronize()      # the library works on its own stream
    total = dom.nelements() if hasattr(dom, "nelements") else len(dom)
    unit = grid_unit(dom.dims) if isinstance(dom, api.CartesianGrid) else 1
    first, count = shard(total, rank, world, unit)
    n, dim = coords_d.shape
    pred = api.fitpredict(model, None, dom, ctx=ctx, first=first, count=count, out=out,
                          _device_samples=(coords_d, values_d, int(n), int(dim), None, [name]), **kw)
    return pred, (first, count)
