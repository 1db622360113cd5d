#!/usr/bin/env python
"""Time the single-process multi-GPU C ABI (gsm_multi_*) on the north-star config and C4: ONE call, N GPUs, fixed
domain (strong scaling), sample upload + NCCL broadcast + bins + prediction + D2H into pinned host memory all inside
the timed region.  usage: multi_bench.py [ns|c4] N [N ...]; one JSON line per N."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
gsm = entry.load_package()
L = gsm._lib

cfg = sys.argv[1] if len(sys.argv) > 1 else "ns"
ns = [int(a) for a in sys.argv[2:]] or [1]
rng = np.random.default_rng(6 if cfg == "ns" else 4)
n = 1_000_000
Xp, zp = gsm.PinnedArray((n, 3), np.float64), gsm.PinnedArray((n,), np.float64)
Xp.array[:] = rng.uniform(0, 256, size=(n, 3))
zp.array[:] = np.sin(Xp.array[:, 0] / 20) + np.cos(Xp.array[:, 2] / 30) + 0.1 * rng.standard_normal(n)
M = 256 ** 3
pm, pv, ps = gsm.PinnedArray((M,), np.float64), gsm.PinnedArray((M,), np.float64), gsm.PinnedArray((M,), np.uint8)
fun = gsm.SphericalVariogram(range=20.0, sill=1.0, nugget=0.1)
model = gsm.OrdinaryKriging(fun) if cfg == "ns" else gsm.UniversalKriging(fun, 2, 3)
tab = gsm.georef({"z": zp.array}, gsm.PointSet(Xp.array))
grid = gsm.CartesianGrid(256, 256, 256)
ref = None
for N in ns:
    devs = list(range(N))
    out = {"mean": pm.array, "var": pv.array, "status": ps.array}
    call = lambda: gsm.fitpredict(model, tab, grid, maxneighbors=32, prob=True, devices=devs, out=out) if N > 1 else \
        gsm.fitpredict(model, tab, grid, maxneighbors=32, prob=True, out=out)
    for _ in range(2):
        call()
    best, rec = 1e30, None
    for _ in range(5):
        t0 = time.perf_counter(); call(); dt = time.perf_counter() - t0
        if dt < best:
            best = dt
            if N > 1:
                mc = gsm.multi_context(devs)
                rec = {"set_samples": None, "predict": mc.timings(), "per_device_total_ms": [mc.device_timings(i)["total_ms"] for i in range(N)]}
            else:
                rec = {"predict": gsm.default_context().timings()}
    mu = np.array(pm.array)
    if ref is None:
        ref = mu
    print(json.dumps({"config": cfg, "n_gpus": N, "e2e_ms": best * 1e3, "predictions_per_s": M / best,
                      "bit_identical_to_first": bool(np.array_equal(mu, ref)), **rec}), flush=True)
