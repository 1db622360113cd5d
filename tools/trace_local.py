#!/usr/bin/env python
"""Phase timeline of the local-Kriging DMMA kernel (developer tool; needs `make -C .../csrc trace`)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
gsm = entry.load_package()
L = gsm._lib
L.lib_path = lambda: os.path.join(ROOT, "geostatsmodels.jl_b200", "lib", "libgsm_b200_trace.so")
lib = L.load()
import torch
tr = torch.zeros(4 * 9 * 40, dtype=torch.int64, device="cuda")
lib.gsm_debug_set_trace2.argtypes = [C.c_void_p]
assert lib.gsm_debug_set_trace2(tr.data_ptr()) == 0
rng = np.random.default_rng(3)
n = 200000
X = rng.uniform(0, 2048, size=(n, 2)); z = rng.standard_normal(n)
ctx = gsm.Context(0)
a = 1 / 2048.0
ctx.set_samples(X * a, z)
v = L.make_variogram(1, 1.0, 0.1, 50 * a); m = L.make_model(1)
t = gsm.Context.grid_targets((0, 0), (a, a), (2048, 512))
om = torch.empty(2048 * 512, dtype=torch.float64, device="cuda"); ov = torch.empty_like(om)
ost = torch.empty(2048 * 512, dtype=torch.uint8, device="cuda")
for _ in range(2):
    tr.zero_()
    ctx.local_krige(v, m, t, 32, out_mean=om, out_var=ov, out_status=ost)
print(ctx.timings())
T = tr.cpu().numpy().reshape(4, 9, 40)
names_sys = {0: "start", 1: "A00 posted", 18: "G handed off", 19: "pool bar2+P5"}
for kb in range(4):
    names_sys.update({2 + 4 * kb: f"col{kb} prepped", 3 + 4 * kb: f"W{kb}", 4 + 4 * kb: f"next diag posted", 5 + 4 * kb: f"col{kb} done"})
for it in range(1, 3):
    base = T[it, :8, 0].min()
    print(f"--- iteration {it} (cycles relative to first warp start)")
    for w in (0, 3, 7):
        print(f" sys warp {w}: " + ", ".join(f"{names_sys[s]}={T[it, w, s] - base}" for s in sorted(names_sys)))
    d = T[it, 8]
    print(" diag warp: " + ", ".join(f"kb{kb}: wait@{d[kb*3]-base} woke@{d[kb*3+1]-base} done@{d[kb*3+2]-base}" for kb in range(4)))
    print(f" group period: {T[it, 0, 0] - T[it - 1, 0, 0]} cycles")
