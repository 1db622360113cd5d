import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as entry
gsm = entry.load_package(); L = gsm._lib
import torch
rng = np.random.default_rng(3); n = 200000
X = rng.uniform(0, 2048, size=(n, 2)); z = rng.standard_normal(n)
ctx = gsm.Context(0); a = 1 / 2048.0
ctx.set_samples(X * a, z)
v = L.make_variogram(1, 1.0, 0.1, 50 * a); m = L.make_model(0, mean=0.0) if hasattr(L, 'make_model') else None
t = gsm.Context.grid_targets((0, 0), (a, a), (2048, 512))
M = 2048 * 512
om = torch.empty(M, dtype=torch.float64, device="cuda"); ov = torch.empty_like(om); ost = torch.empty(M, dtype=torch.uint8, device="cuda")
for k in (48, 64):
    for _ in range(2):
        ctx.local_krige(v, L.make_model(1), t, k, out_mean=om, out_var=ov, out_status=ost)
    print(k, {kk: round(x, 2) for kk, x in ctx.timings().items() if kk in ("knn_ms", "solve_ms")})
