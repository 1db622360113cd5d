#!/usr/bin/env python
"""k = 64 timing (C5 column): SimpleKriging, 1 M 2-D samples -> 2048^2."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
gsm = entry.load_package()
k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(5)
X = rng.uniform(0, 2048, size=(1_000_000, 2)); z = rng.standard_normal(1_000_000)
tab = gsm.georef({"z": z}, gsm.PointSet(X)); grid = gsm.CartesianGrid(2048, 1024)
sk = gsm.SimpleKriging(gsm.SphericalVariogram(range=2048 / 20, nugget=0.1), float(z.mean()))
ctx = gsm.Context(0)
for _ in range(3):
    p = gsm.fitpredict(sk, tab, grid, maxneighbors=k, prob=True, ctx=ctx).z
    t = ctx.timings()
print(f"k={k} knn {t['knn_ms']:.2f} solve {t['solve_ms']:.2f} ms for {grid.nelements()} targets; per 2^20: {t['solve_ms'] / grid.nelements() * (1 << 20):.2f} ms; nan {int(np.isnan(np.asarray(p.mu)).sum())}")
