#!/usr/bin/env python
"""Randomised parity sweep of the global (all-samples) Kriging / IDW path against the numpy oracle."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
gsm = entry.load_package()
o = entry.load_oracle()


def one(seed):
    rng = np.random.default_rng(seed)
    dim = int(rng.choice([2, 2, 3]))
    kind = str(rng.choice(["sk", "ok", "ok", "uk", "idw"]))
    vk = int(rng.choice([0, 1, 2]))
    n = int(rng.choice([5, 17, 64, 127, 128, 129, 300, 513, 700]))
    M = int(rng.choice([1, 63, 64, 65, 1000, 8191, 8193, 12000]))
    X = rng.uniform(0, 100, size=(n, dim))
    z = np.sin(X[:, 0] / 15) + 0.1 * rng.standard_normal(n)
    P = rng.uniform(-5, 105, size=(M, dim))
    vario = o.Variogram(vk, 1.0, 0.1 if vk == 0 else float(rng.choice([0.0, 0.1])), float(rng.uniform(10, 60)))
    deg = int(rng.choice([1, 2])) if kind == "uk" else None
    if kind == "uk":   # fewer samples than monomials is singular by construction
        n = max(n, {(2, 1): 3, (2, 2): 6, (3, 1): 4, (3, 2): 10}[(dim, deg)] + 7)
        X = rng.uniform(0, 100, size=(n, dim))
        z = np.sin(X[:, 0] / 15) + 0.1 * rng.standard_normal(n)
    desc = f"seed {seed}: {kind} vk={vk} dim={dim} n={n} M={M} deg={deg}"
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    if kind == "idw":
        e = float(rng.choice([1, 2, 3, 2.5]))
        mu, _, _ = o.fitpredict(o.IDWModel(e), X, z, P, prob=False, neighbors=False)
        got = gsm.fitpredict(gsm.IDW(e), tab, gsm.PointSet(P), neighbors=False).z
        return bool(np.allclose(got, mu, rtol=1e-10, atol=1e-12)), desc
    om = {"sk": lambda: o.simple_kriging(vario, float(z.mean())), "ok": lambda: o.ordinary_kriging(vario),
          "uk": lambda: o.universal_kriging(vario, deg, dim)}[kind]()
    cls = {0: gsm.GaussianVariogram, 1: gsm.SphericalVariogram, 2: gsm.ExponentialVariogram}[vk]
    fun = cls(sill=vario.sill, range=vario.range, nugget=vario.nugget)
    gm = {"sk": lambda: gsm.SimpleKriging(fun, float(z.mean())), "ok": lambda: gsm.OrdinaryKriging(fun),
          "uk": lambda: gsm.UniversalKriging(fun, deg, dim)}[kind]()
    mu, var, st = o.fitpredict(om, X, z, P, prob=True, neighbors=False)
    try:
        got = gsm.fitpredict(gm, tab, gsm.PointSet(P), neighbors=False, prob=True).z
    except ValueError as ex:
        return bool((var <= 0).any()), desc + f" [{ex}; oracle min var {var.min():.3e}]"
    tol = 1e-9 if kind != "uk" else 1e-6
    ok = np.allclose(got.mu, mu, rtol=tol, atol=tol) and np.allclose(got.sigma, var, rtol=tol, atol=tol)
    if not ok:
        desc += " maxerr mu %.2e var %.2e" % (np.abs(got.mu - mu).max(), np.abs(got.sigma - var).max())
    return bool(ok), desc


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=None)
    args = ap.parse_args()
    seeds = [args.seed] if args.seed is not None else list(range(3000, 3000 + args.cases))
    t0 = time.time(); nfail = 0
    for sd in seeds:
        try:
            ok, desc = one(sd)
        except Exception as ex:
            ok, desc = False, f"seed {sd}: EXCEPTION {ex!r}"
        if not ok:
            nfail += 1
            print("FAIL", desc, flush=True)
    print(f"{len(seeds) - nfail} of {len(seeds)} global cases agree with the oracle ({time.time() - t0:.0f} s)")
