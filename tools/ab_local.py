#!/usr/bin/env python
"""A/B the local-Kriging kernels (GSM_LOCAL_IMPL) on the C3 workload: timings and element-wise differences."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
gsm = entry.load_package(); L = gsm._lib
import torch
ny = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
rng = np.random.default_rng(3)
X = rng.uniform(0, 2048, size=(n, 2)); z = rng.standard_normal(n)
ctx = gsm.Context(0); a = float(sys.argv[3]) if len(sys.argv) > 3 else 1 / 2048.0
ctx.set_samples(X * a, z)
v = L.make_variogram(1, 1.0, 0.1, 50 * a); m = L.make_model(1)
t = gsm.Context.grid_targets((0, 0), (a, a), (2048, ny))
M = 2048 * ny
res = {}
for impl in ("pipe", "shr"):
    os.environ["GSM_LOCAL_IMPL"] = impl
    om = torch.empty(M, dtype=torch.float64, device="cuda"); ov = torch.empty_like(om); ost = torch.empty(M, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ctx.local_krige(v, m, t, 32, out_mean=om, out_var=ov, out_status=ost)
    print(impl, {k: round(x, 3) for k, x in ctx.timings().items() if k in ("knn_ms", "solve_ms")})
    res[impl] = (om.cpu().numpy(), ov.cpu().numpy(), ost.cpu().numpy())
a_, b_ = res["pipe"], res["shr"]
dm = np.abs(a_[0] - b_[0]) / np.maximum(np.abs(a_[0]), 1e-9); dv = np.abs(a_[1] - b_[1]) / np.maximum(np.abs(a_[1]), 1e-12)
bad = np.flatnonzero(~(dm < 1e-9) | ~(dv < 1e-9) | (a_[2] != b_[2]))
print("max rel diff mean %.3e var %.3e; mismatches %d of %d; status!=0: pipe %d shr %d" % (np.nanmax(dm), np.nanmax(dv), bad.size, M, (a_[2] != 0).sum(), (b_[2] != 0).sum()))
print("status values shr:", np.unique(b_[2], return_counts=True))
if bad.size:
    print("first bad:", bad[:24]); print("x:", bad[:24] % 2048, "y:", bad[:24] // 2048)
    print("shr mean/var at bad:", b_[0][bad[:6]], b_[1][bad[:6]], "pipe:", a_[0][bad[:6]], a_[1][bad[:6]])
    ys = np.unique(bad // 2048); print("bad rows:", ys[:40], "count rows", ys.size)
