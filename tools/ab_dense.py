#!/usr/bin/env python
"""A/B the local-Kriging kernels when the samples are as dense as / denser than the target grid (large pools)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
gsm = entry.load_package(); L = gsm._lib
import torch
g = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
rng = np.random.default_rng(3)
X = rng.uniform(0, g, size=(n, 2)); z = rng.standard_normal(n)
ctx = gsm.Context(0); a = 1.0 / g
ctx.set_samples(X * a, z)
v = L.make_variogram(1, 1.0, 0.1, 20 * a); m = L.make_model(1)
t = gsm.Context.grid_targets((0, 0), (a, a), (g, g))
M = g * g
res = {}
for impl in ("pipe", "shr"):
    os.environ["GSM_LOCAL_IMPL"] = impl
    om = torch.empty(M, dtype=torch.float64, device="cuda"); ov = torch.empty_like(om); ost = torch.empty(M, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ctx.local_krige(v, m, t, 32, out_mean=om, out_var=ov, out_status=ost)
    print(impl, {k: round(x, 3) for k, x in ctx.timings().items() if k in ("knn_ms", "solve_ms", "index_ms")})
    res[impl] = (om.cpu().numpy(), ov.cpu().numpy())
print("max abs diff mean %.2e var %.2e" % (np.abs(res["pipe"][0] - res["shr"][0]).max(), np.abs(res["pipe"][1] - res["shr"][1]).max()))
