import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
import __graft_entry__ as entry
gsm = entry.load_package()
for cfg in ("c3", "ns"):
    if cfg == "c3":
        rng = np.random.default_rng(3); n = 200_000
        X = rng.uniform(0, 2048, size=(n, 2)); z = rng.standard_normal(n); grid = gsm.CartesianGrid(2048, 2048)
        model = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=50.0, nugget=0.1))
    else:
        rng = np.random.default_rng(6); n = 1_000_000
        X = rng.uniform(0, 256, size=(n, 3)); z = rng.standard_normal(n); grid = gsm.CartesianGrid(256, 256, 256)
        model = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=20.0, nugget=0.1))
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    for occ in (1.0, 1.5, 2.0, 2.5, 3.0, 4.0, 6.0):
        ctx = gsm.Context(0); ctx.set_option("bin_occupancy", occ)
        best = 1e9
        for _ in range(3):
            gsm.fitpredict(model, tab, grid, maxneighbors=32, prob=True, ctx=ctx); best = min(best, ctx.timings()["knn_ms"])
        print(cfg, occ, round(best, 2), flush=True)
        ctx.close()
