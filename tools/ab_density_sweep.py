#!/usr/bin/env python
"""Sweep the sample density per target cell and compare pipe / shr / default dispatch (2-D or 3-D)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
gsm = entry.load_package(); L = gsm._lib
import torch
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = 1024 if dim == 2 else 128
M = g ** dim
ctx = gsm.Context(0); a = 1.0 / g
rng = np.random.default_rng(3)
for s in ([0.03, 0.1, 0.2, 0.35, 0.5, 1.0] if dim == 2 else [0.02, 0.06, 0.15, 0.3, 0.6, 1.5]):
    n = int(s * M)
    X = rng.uniform(0, g, size=(n, dim)); z = rng.standard_normal(n)
    ctx.set_samples(X * a, z)
    v = L.make_variogram(1, 1.0, 0.1, 20 * a); m = L.make_model(1)
    t = gsm.Context.grid_targets((0,) * dim, (a,) * dim, (g,) * dim)
    om = torch.empty(M, dtype=torch.float64, device="cuda"); ov = torch.empty_like(om); ost = torch.empty(M, dtype=torch.uint8, device="cuda")
    out = []
    for impl in ("pipe", "shr", "auto"):
        ctx.set_option("local_impl", impl)
        for _ in range(3):
            ctx.local_krige(v, m, t, 32, out_mean=om, out_var=ov, out_status=ost)
        out.append(round(ctx.timings()["solve_ms"], 2))
    rk = (32 / (np.pi * s)) ** 0.5 if dim == 2 else (3 * 32 / (4 * np.pi * s)) ** (1 / 3)
    print(f"dim {dim} samples/cell {s:5.2f} r_k {rk:5.1f} cells: pipe {out[0]:7.2f} shr {out[1]:7.2f} default {out[2]:7.2f} ms")
