import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
gsm = entry.load_package(); L = gsm._lib
name = sys.argv[1]
L.lib_path = lambda: os.path.join(ROOT, "geostatsmodels.jl_b200", "lib", name)
import torch
rng = np.random.default_rng(3); n = 200000
X = rng.uniform(0, 2048, size=(n, 2)); z = rng.standard_normal(n)
ctx = gsm.Context(0); a = 1 / 2048.0
ctx.set_samples(X * a, z)
v = L.make_variogram(1, 1.0, 0.1, 50 * a); m = L.make_model(1)
t = gsm.Context.grid_targets((0, 0), (a, a), (2048, 2048))
M = 2048 * 2048
om = torch.empty(M, dtype=torch.float64, device="cuda"); ov = torch.empty_like(om); ost = torch.empty(M, dtype=torch.uint8, device="cuda")
for _ in range(3):
    ctx.local_krige(v, m, t, 32, out_mean=om, out_var=ov, out_status=ost)
print(name, {k: round(x, 3) for k, x in ctx.timings().items() if k in ("knn_ms", "solve_ms")})
