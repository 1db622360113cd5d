"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes replicate the samples with
one broadcast, split the target range like the reference's thread chunks, and the shards (predicted here
by the oracle standing in for the GPU) concatenate to the single-rank answer."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_partition_matches_index_chunks(gsm):
    sh = gsm.distributed.shard
    for total in (0, 1, 7, 100, 4194304, 16777216 + 3):
        for world in (1, 2, 3, 8):
            parts = [sh(total, r, world) for r in range(world)]
            assert parts[0][0] == 0
            assert sum(c for _, c in parts) == total
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1                      # contiguous
            sizes = [c for _, c in parts]
            assert max(sizes) - min(sizes) <= 1           # balanced like index_chunks(n=nthreads)
            assert sizes == sorted(sizes, reverse=True)   # the larger chunks come first


def test_shard_units_keep_patch_boundaries(gsm):
    sh, gu = gsm.distributed.shard, gsm.distributed.grid_unit
    assert gu((256, 256, 256)) == 2 * 256 * 256 and gu((2048, 2048)) == 2 * 2048 and gu((17,)) == 1
    for dims in ((256, 256, 256), (2048, 2048), (33, 21, 9), (100, 7)):
        total, unit = int(np.prod(dims)), gu(dims)
        for world in (1, 2, 3, 4, 8):
            parts = [sh(total, r, world, unit) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1 and f1 % unit == 0


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    gsm = entry.load_package()
    o = entry.load_oracle()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if rank == 0:
        rng = np.random.default_rng(11)
        X = rng.uniform(0, 50, size=(400, 2))
        z = rng.standard_normal(400)
    else:
        X = z = None
    c, v = gsm.distributed.replicate_samples(X, z, src=0)
    Xr, zr = c.numpy(), v.numpy()
    grid = o.Grid((0.0, 0.0), (2.5, 2.5), (20, 17))
    first, count = gsm.distributed.shard(grid.nelements, rank, world)
    model = o.ordinary_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.1, 15.0))
    mu, var, st = o.fitpredict(model, Xr, zr, grid, maxneighbors=12, prob=True, first=first, count=count)
    np.savez(os.path.join(tmp, f"part{rank}.npz"), mu=mu, var=var, first=first, count=count, X=Xr, z=zr)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_broadcast_and_shards(tmp_path, oracle):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"part{r}.npz") for r in range(world)]
    assert np.array_equal(parts[0]["X"], parts[1]["X"]) and np.array_equal(parts[0]["z"], parts[1]["z"])
    o = oracle
    grid = o.Grid((0.0, 0.0), (2.5, 2.5), (20, 17))
    model = o.ordinary_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.1, 15.0))
    mu, var, _ = o.fitpredict(model, parts[0]["X"], parts[0]["z"], grid, maxneighbors=12, prob=True)
    assert np.array_equal(np.concatenate([p["mu"] for p in parts]), mu)
    assert np.array_equal(np.concatenate([p["var"] for p in parts]), var)
    assert int(parts[0]["count"]) + int(parts[1]["count"]) == grid.nelements
