import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
gsm = entry.load_package(); L = gsm._lib
rng = np.random.default_rng(3); n = 200000
X = rng.uniform(0, 2048, size=(n, 2)); z = rng.standard_normal(n)
ctx = gsm.Context(0)
M = 2048 * 2048
Xh = gsm.PinnedArray((n, 2), np.float64); zh = gsm.PinnedArray((n,), np.float64)
Xh.array[:] = X; zh.array[:] = z
pm, pv, ps = gsm.PinnedArray((M,), np.float64), gsm.PinnedArray((M,), np.float64), gsm.PinnedArray((M,), np.uint8)
table = gsm.georef({"z": zh.array}, gsm.PointSet(Xh.array))
grid = gsm.CartesianGrid(2048, 2048)
okm = gsm.OrdinaryKriging(gsm.SphericalVariogram(sill=1.0, range=50.0, nugget=0.1))
for i in range(6):
    t0 = time.perf_counter()
    gsm.fitpredict(okm, table, grid, maxneighbors=32, prob=True, ctx=ctx, out={"mean": pm.array, "var": pv.array, "status": ps.array})
    t1 = time.perf_counter()
    print("fitpredict wall %.2f ms; lib local_krige timings:" % ((t1 - t0) * 1e3), {k: round(v, 2) for k, v in ctx.timings().items()})
# pieces
a = 1 / 2048.0
t0 = time.perf_counter(); ctx.set_samples(Xh.array * a, zh.array); t1 = time.perf_counter()
print("set_samples wall %.2f ms" % ((t1 - t0) * 1e3), ctx.timings())
v = L.make_variogram(1, 1.0, 0.1, 50 * a); m = L.make_model(1)
t = gsm.Context.grid_targets((0, 0), (a, a), (2048, 2048))
t0 = time.perf_counter(); ctx.local_krige(v, m, t, 32, out_mean=pm.array, out_var=pv.array, out_status=ps.array); t1 = time.perf_counter()
print("local_krige wall %.2f ms" % ((t1 - t0) * 1e3), ctx.timings())
