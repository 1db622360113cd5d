#!/usr/bin/env python
"""Phase timeline of the shared-factor local-Kriging kernel (developer tool; needs `make -C .../csrc trace`)."""
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
tr = torch.zeros(4 * 10 * 16, dtype=torch.int64, device="cuda")
lib.gsm_debug_set_trace3.argtypes = [C.c_void_p]
assert lib.gsm_debug_set_trace3(tr.data_ptr()) == 0
rng = np.random.default_rng(3)
ctx = gsm.Context(0)
if len(sys.argv) > 1 and sys.argv[1] == "3d":   # needs `make trace EXTRA="-DGSM_TRACE -DGSM_TRACE_DIM3=1"`
    n = 1000000
    X = rng.uniform(0, 256, size=(n, 3)); z = rng.standard_normal(n)
    a = 1 / 256.0
    ctx.set_samples(X * a, z)
    v = L.make_variogram(1, 1.0, 0.1, 20 * a); m = L.make_model(1)
    t = gsm.Context.grid_targets((0, 0, 0), (a, a, a), (256, 256, 16))
    M = 256 * 256 * 16
else:
    n = 200000
    X = rng.uniform(0, 2048, size=(n, 2)); z = rng.standard_normal(n)
    a = 1 / 2048.0
    ctx.set_samples(X * a, z)
    v = L.make_variogram(1, 1.0, 0.1, 50 * a); m = L.make_model(1)
    t = gsm.Context.grid_targets((0, 0), (a, a), (2048, 512))
    M = 2048 * 512
om = torch.empty(M, dtype=torch.float64, device="cuda"); ov = torch.empty_like(om)
ost = torch.empty(M, dtype=torch.uint8, device="cuda")
for _ in range(2):
    tr.zero_()
    ctx.local_krige(v, m, t, 32, out_mean=om, out_var=ov, out_status=ost)
print(ctx.timings())
T = tr.cpu().numpy().reshape(4, 10, 16)
base = T[0, 0, 0]
ns = ["start", "prepped", "-", "slot", "landed", "Q1", "Q4", "fin", "endbar"]
nd = ["woke", "done", "epi-done"]
nf = ["start", "hashed", "ranked", "FFREE", "published", "factored"]
for it in range(4):
    print(f"--- job {20 + it}")
    for w in (0, 5):
        print(f"  sys{w}: " + ", ".join(f"{ns[s]}={T[it, w, s] - base}" for s in range(9) if s != 2))
    print("  diag: " + ", ".join(f"{nd[s]}={T[it, 8, s] - base}" for s in range(3)))
