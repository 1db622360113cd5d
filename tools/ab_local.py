#!/usr/bin/env python
"""A/B the local-Kriging kernels (gsm_set_option local_impl) on a BASELINE workload: timings and differences.
usage: ab_local.py [c3|ns|c4] impl [impl ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
gsm = entry.load_package()

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
impls = sys.argv[2:] or ["shr", "patch"]
if cfg == "c3":
    rng = np.random.default_rng(3); n = 200_000
    X = rng.uniform(0, 2048, size=(n, 2)); z = np.sin(X[:, 0] / 20) + np.cos(X[:, 1] / 30) + 0.1 * rng.standard_normal(n)
    grid = gsm.CartesianGrid(2048, 2048)
    model = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=50.0, sill=1.0, nugget=0.1))
else:
    rng = np.random.default_rng(6 if cfg == "ns" else 4); n = 1_000_000
    X = rng.uniform(0, 256, size=(n, 3)); z = np.sin(X[:, 0] / 20) + np.cos(X[:, 2] / 30) + 0.1 * rng.standard_normal(n)
    grid = gsm.CartesianGrid(256, 256, 256)
    fun = gsm.SphericalVariogram(range=20.0, sill=1.0, nugget=0.1)
    model = gsm.OrdinaryKriging(fun) if cfg == "ns" else gsm.UniversalKriging(fun, 2, 3)
tab = gsm.georef({"z": z}, gsm.PointSet(X))
ref = None
for spec in impls:
    impl, _, extra = spec.partition(":")
    ctx = gsm.Context(0)
    ctx.set_option("local_impl", impl)
    for kv in filter(None, extra.split(",")):
        key, _, val = kv.partition("=")
        ctx.set_option(key, val)
    best = None
    for _ in range(3):
        p = gsm.fitpredict(model, tab, grid, maxneighbors=32, prob=True, ctx=ctx).z
        t = ctx.timings()
        if best is None or t["solve_ms"] < best["solve_ms"]:
            best = t
    mu, var = np.array(p.mu), np.array(p.sigma)
    line = f"{cfg} {spec:24s} knn {best['knn_ms']:7.2f} ms  index {best['index_ms']:6.2f} ms  solve {best['solve_ms']:8.2f} ms  nan {int(np.isnan(mu).sum())}"
    if ref is not None:
        line += f"  max|dmu| {np.nanmax(np.abs(mu - ref[0])):.2e}  max|dvar| {np.nanmax(np.abs(var - ref[1])):.2e}"
    else:
        ref = (mu, var)
    print(line, flush=True)
    ctx.close()
