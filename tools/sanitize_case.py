#!/usr/bin/env python
"""Small steady-state run of the local-Kriging path for compute-sanitizer (memcheck / racecheck)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
gsm = entry.load_package()
dims = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (160, 48)
rng = np.random.default_rng(5)
n = 3000
X = rng.uniform(0, 1, size=(n, 2)) * np.array(dims, dtype=float)
z = rng.standard_normal(n)
ok = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=25.0, nugget=0.1))
pred = gsm.fitpredict(ok, gsm.georef({"z": z}, gsm.PointSet(X)), gsm.CartesianGrid(*dims), maxneighbors=32, prob=True)
print("ok", float(pred.z.mu.mean()), float(pred.z.sigma.mean()))
