#!/usr/bin/env python
"""Time (and spot-check against the oracle) the BASELINE.json configs C1..C5 on one GPU.
Developer / documentation tool: prints one JSON line per config.  `--quick` shrinks the domains."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--only", default="")
ap.add_argument("--gpus", type=int, default=1)
args = ap.parse_args()
gsm = entry.load_package()
o = entry.load_oracle()
import gsm_oracle_c as oc
L = gsm._lib
ctx = gsm.Context(0)
Q = 4 if args.quick else 1


def check(name, got_mu, got_var, model, X, z, grid, nsub, **kw):
    """parity on a strided subsample of the domain (C oracle for local, numpy oracle for global/IDW)"""
    M = grid.nelements
    idx = np.unique(np.linspace(0, M - 1, nsub).astype(np.int64))
    T = grid.centroids()[idx] if M <= 5_000_000 else np.stack([grid.centroids(int(i), 1)[0] for i in idx])
    if kw.get("neighbors", True):
        mu, var, st, _, _ = oc.fitpredict(model, X, z, _PointsDom(T, grid, X), prob=True, maxneighbors=kw["k"], nthreads=8)
    else:
        mu, var, st = o.fitpredict(model, X, z, _PointsDom(T, grid, X), prob=True, neighbors=False)
    emu = np.max(np.abs(got_mu[idx] - mu) / np.maximum(np.abs(mu), 1e-12))
    out = {"max_rel_err_mean": float(emu)}
    if got_var is not None and var is not None:
        out["max_rel_err_var"] = float(np.max(np.abs(got_var[idx] - var) / np.maximum(np.abs(var), 1e-12)))
    return out


class _PointsDom(np.ndarray):
    """Explicit target points that still scale like the full grid (alpha from the grid's bounding box)."""
    def __new__(cls, T, grid, X):
        obj = np.asarray(T).view(cls)
        obj._grid = grid
        return obj


_orig_absmax = o._dom_absmax
o._dom_absmax = lambda dom: dom._grid.bbox_absmax() if isinstance(dom, _PointsDom) else _orig_absmax(dom)


def timed(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); best = min(best, time.perf_counter() - t0)
    return r, best


def run(name, fn):
    if args.only and args.only not in name:
        return
    try:
        print(json.dumps({"config": name, **fn()}), flush=True)
    except Exception as e:  # keep going: this is a survey tool
        print(json.dumps({"config": name, "error": repr(e)}), flush=True)


def c1():
    rng = np.random.default_rng(1); n = 10_000
    X = rng.uniform(0, 256, size=(n, 2)); z = np.sin(X[:, 0] / 20) + np.cos(X[:, 1] / 30) + 0.1 * rng.standard_normal(n)
    grid = gsm.CartesianGrid(256 // Q, 256 // Q)
    pred, dt = timed(lambda: gsm.fitpredict(gsm.IDW(2), gsm.georef({"z": z}, gsm.PointSet(X)), grid, neighbors=False, ctx=ctx))
    og = o.Grid((0.0, 0.0), (1.0, 1.0), grid.dims)
    return {"M": grid.nelements(), "e2e_ms": dt * 1e3, "kernel_ms": ctx.timings()["solve_ms"],
            **check("c1", pred.z, None, o.IDWModel(2), X, z, og, 200, neighbors=False)}


def c2():
    rng = np.random.default_rng(2); n = 4000 // Q
    X = rng.uniform(0, 1024, size=(n, 2)); z = np.sin(X[:, 0] / 80) + np.cos(X[:, 1] / 120) + 0.1 * rng.standard_normal(n)
    grid = gsm.CartesianGrid(1024 // Q, 1024 // Q)
    model = gsm.OrdinaryKriging(gsm.GaussianVariogram(range=100.0, sill=1.0, nugget=0.01))
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    t0 = time.perf_counter(); pred = gsm.fitpredict(model, tab, grid, neighbors=False, prob=True, ctx=ctx); dt = time.perf_counter() - t0
    og = o.Grid((0.0, 0.0), (1.0, 1.0), grid.dims)
    om = o.ordinary_kriging(o.Variogram(o.GAUSSIAN, 1.0, 0.01, 100.0))
    return {"N": n, "M": grid.nelements(), "e2e_ms_incl_fit": dt * 1e3, "predict_kernel_ms": ctx.timings()["solve_ms"],
            **check("c2", pred.z.mu, pred.z.sigma, om, X, z, og, 60, neighbors=False)}


def c3():
    rng = np.random.default_rng(3); n = 200_000 // (Q * Q)
    X = rng.uniform(0, 2048 // Q, size=(n, 2)); z = np.sin(X[:, 0] / 20) + np.cos(X[:, 1] / 30) + 0.1 * rng.standard_normal(n)
    grid = gsm.CartesianGrid(2048 // Q, 2048 // Q)
    model = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=50.0, sill=1.0, nugget=0.1))
    pred, dt = timed(lambda: gsm.fitpredict(model, gsm.georef({"z": z}, gsm.PointSet(X)), grid, maxneighbors=32, prob=True, ctx=ctx))
    t = ctx.timings()
    om = o.ordinary_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.1, 50.0))
    return {"M": grid.nelements(), "e2e_ms": dt * 1e3, "knn_ms": t["knn_ms"], "solve_ms": t["solve_ms"],
            **check("c3", pred.z.mu, pred.z.sigma, om, X, z, o.Grid((0.0, 0.0), (1.0, 1.0), grid.dims), 2000, k=32)}


def c4():
    rng = np.random.default_rng(4); n = 1_000_000 // (Q ** 3)
    X = rng.uniform(0, 256 // Q, size=(n, 3)); z = np.sin(X[:, 0] / 20) + np.cos(X[:, 2] / 30) + 0.1 * rng.standard_normal(n)
    grid = gsm.CartesianGrid(256 // Q, 256 // Q, 256 // Q)
    model = gsm.UniversalKriging(gsm.SphericalVariogram(range=20.0, sill=1.0, nugget=0.1), 2, 3)
    pred, dt = timed(lambda: gsm.fitpredict(model, gsm.georef({"z": z}, gsm.PointSet(X)), grid, maxneighbors=32, prob=True, ctx=ctx), reps=1)
    t = ctx.timings()
    om = o.universal_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.1, 20.0), 2, 3)
    return {"N": n, "M": grid.nelements(), "e2e_ms": dt * 1e3, "knn_ms": t["knn_ms"], "solve_ms": t["solve_ms"],
            **check("c4", pred.z.mu, pred.z.sigma, om, X, z, o.Grid((0.0,) * 3, (1.0,) * 3, grid.dims), 1000, k=32)}


def ns():
    """north-star target: local OK k=32 on 1M samples -> 256^3 grid"""
    rng = np.random.default_rng(6); n = 1_000_000 // (Q ** 3)
    X = rng.uniform(0, 256 // Q, size=(n, 3)); z = np.sin(X[:, 0] / 20) + np.cos(X[:, 2] / 30) + 0.1 * rng.standard_normal(n)
    grid = gsm.CartesianGrid(256 // Q, 256 // Q, 256 // Q)
    model = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=20.0, sill=1.0, nugget=0.1))
    pred, dt = timed(lambda: gsm.fitpredict(model, gsm.georef({"z": z}, gsm.PointSet(X)), grid, maxneighbors=32, prob=True, ctx=ctx), reps=2)
    t = ctx.timings()
    om = o.ordinary_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.1, 20.0))
    return {"N": n, "M": grid.nelements(), "e2e_ms": dt * 1e3, "knn_ms": t["knn_ms"], "solve_ms": t["solve_ms"],
            **check("ns", pred.z.mu, pred.z.sigma, om, X, z, o.Grid((0.0,) * 3, (1.0,) * 3, grid.dims), 1000, k=32)}


def c5():
    """C5: SimpleKriging + IDW sweep, samples 10^4 .. 10^7, k = 8 .. 64, 2-D (2048^2) and 3-D (256^3) grids, random and
    LATTICE samples (integer lattice: distance ties everywhere), on --gpus GPUs through fitpredict(devices=...)."""
    out = {}
    devs = list(range(args.gpus)) if args.gpus > 1 else None
    tctx = None if devs else ctx
    def tim():
        if devs:
            mc = gsm.multi_context(devs)
            t = [mc.device_timings(i) for i in range(len(devs))]
            return {"knn_ms": max(x["knn_ms"] for x in t), "solve_ms": max(x["solve_ms"] for x in t)}
        return ctx.timings()
    for dim, dims in ((2, (2048 // Q, 2048 // Q)), (3, (256 // Q,) * 3)):
        for n in (10_000, 100_000, 1_000_000, 10_000_000):
            if args.quick and n > 1_000_000:
                continue
            for lattice in (False, True):
                if lattice and n != 1_000_000:
                    continue
                rng = np.random.default_rng(5)
                X = rng.uniform(0, dims[0], size=(n, dim))
                if lattice:   # samples on a lattice of pitch ~ (volume / n)^(1/dim): every distance is tied many times
                    pitch = dims[0] / round(n ** (1.0 / dim))
                    X = (np.floor(X / pitch) + 0.5) * pitch
                    X = np.unique(X, axis=0)
                z = rng.standard_normal(X.shape[0])
                tab = gsm.georef({"z": z}, gsm.PointSet(X)); grid = gsm.CartesianGrid(*dims)
                tag = f"{dim}D N={X.shape[0]}{' lattice' if lattice else ''}"
                for k in (8, 16, 32, 64):
                    sk = gsm.SimpleKriging(gsm.SphericalVariogram(range=dims[0] / 20, nugget=0.1), float(z.mean()))
                    _, dt = timed(lambda: gsm.fitpredict(sk, tab, grid, maxneighbors=k, prob=True, ctx=tctx, devices=devs), reps=1)
                    t = tim()
                    out[f"SK {tag} k={k}"] = {"e2e_ms": round(dt * 1e3, 1), "knn_ms": round(t["knn_ms"], 2), "solve_ms": round(t["solve_ms"], 2)}
                for k in (8, 32, 64):
                    _, dt = timed(lambda: gsm.fitpredict(gsm.IDW(2), tab, grid, maxneighbors=k, ctx=tctx, devices=devs), reps=1)
                    t = tim()
                    out[f"IDW {tag} k={k}"] = {"e2e_ms": round(dt * 1e3, 1), "knn_ms": round(t["knn_ms"], 2), "solve_ms": round(t["solve_ms"], 2)}
    return {"gpus": args.gpus, "M2d": (2048 // Q) ** 2, "M3d": (256 // Q) ** 3, "results": out}


for name, fn in (("C1 IDW global", c1), ("C2 OK Gaussian global", c2), ("C3 local OK k=32", c3), ("C4 UK quadratic 3D", c4), ("NS OK 3D 1M->256^3", ns), ("C5 sweep", c5)):
    run(name, fn)
