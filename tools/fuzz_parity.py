#!/usr/bin/env python
"""Randomised parity sweep of the neighbourhood path against the C oracle (developer tool, needs a GPU).
Every case draws model / variogram / dimension / k / grid shape / slab / radius / minneighbors at random; failures
print the seed so they can be replayed with --seed."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

gsm = entry.load_package()
o = entry.load_oracle()
import gsm_oracle_c as oc


def one(seed):
    rng = np.random.default_rng(seed)
    dim = int(rng.choice([2, 2, 3]))
    kind = str(rng.choice(["sk", "ok", "ok", "uk", "idw"]))
    vk = int(rng.choice([0, 1, 1, 2]))
    k = int(rng.choice([3, 8, 9, 12, 16, 17, 24, 31, 32, 33, 40, 48, 64]))
    if dim == 2:
        dims = (int(rng.integers(1, 200)), int(rng.integers(1, 120)))
    else:
        dims = (int(rng.integers(1, 40)), int(rng.integers(1, 40)), int(rng.integers(1, 30)))
    M = int(np.prod(dims))
    ext = np.array([float(d) for d in dims]) * float(rng.choice([0.5, 1.0, 3.0]))
    n = int(rng.integers(max(k + 5, 70), 20000))
    X = rng.uniform(0, 1, size=(n, dim)) * ext
    dups = rng.random() < 0.2
    if dups:   # exact duplicates of target centroids / lattice samples (distance ties, singular systems)
        X[: n // 10] = np.floor(X[: n // 10]) + 0.5
    z = np.sin(X[:, 0] / 7) + 0.1 * rng.standard_normal(n)
    first = int(rng.integers(0, M)) if rng.random() < 0.4 else 0
    count = int(rng.integers(1, M - first + 1)) if rng.random() < 0.4 else None
    nbh = float(rng.uniform(2.0, 12.0)) if rng.random() < 0.3 else None
    minn = int(rng.integers(1, 6)) if nbh is not None else 1
    rngv = float(rng.uniform(3.0, 40.0))
    # the sill spans 26 orders of magnitude (pivot tests must be relative to it); Gaussian keeps sill >> its 1e-6 floor
    sill = float(rng.choice([1.0, 1.0, 7.3, 1e-4, 1e13, 1e-13])) if vk != 0 else float(rng.choice([1.0, 1.0, 7.3, 1e4]))
    vario = o.Variogram(vk, sill, sill * (float(rng.choice([0.0, 0.05, 0.2])) if vk != 0 else 0.1), rngv)
    deg = int(rng.choice([1, 2])) if kind == "uk" else None
    if kind == "uk":   # keep the drift system over-determined: fewer points than monomials is singular by construction
        ncon = {(2, 1): 3, (2, 2): 6, (3, 1): 4, (3, 2): 10}[(dim, deg)]
        k = max(k, ncon + 4)
        minn = max(minn, ncon + 3)
    og = o.Grid(tuple(0.0 for _ in dims), tuple(float(e / d) for e, d in zip(ext, dims)), dims)
    grid = gsm.CartesianGrid(tuple(0.0 for _ in dims), tuple(float(e) for e in ext), dims=dims)
    if rng.random() < 0.25:   # explicit points instead of a grid (the per-system kernels; the order carries no locality)
        P = rng.uniform(-0.05, 1.05, size=(min(M, 4000), dim)) * ext
        og, grid = P, gsm.PointSet(P)
        first = min(first, P.shape[0] - 1)
        count = None if count is None else max(1, min(count, P.shape[0] - first))
        dims = ("points", P.shape[0])
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    kw = dict(maxneighbors=k, minneighbors=minn, first=first)
    desc = f"seed {seed}: {kind} vk={vk} dim={dim} k={k} dims={dims} n={n} first={first} count={count} nbh={nbh} minn={minn} deg={deg}"
    if kind == "idw":
        e = float(rng.choice([1, 2, 3, 2.5]))
        mu, _, st, _, _ = oc.fitpredict(o.IDWModel(e), X, z, og, prob=False, neighborhood=nbh, count=count, nthreads=8, **kw)
        got = gsm.fitpredict(gsm.IDW(e), tab, grid, neighborhood=nbh, count=-1 if count is None else count, **kw).z
        miss = np.ma.getmaskarray(got) if np.ma.isMaskedArray(got) else np.zeros(got.shape, bool)
        ok = np.array_equal(miss, st == 1) and np.allclose(np.ma.filled(got, 0.0)[~miss], mu[~miss], rtol=1e-10, atol=1e-12)
        return ok, desc
    om = {"sk": lambda: o.simple_kriging(vario, float(z.mean())), "ok": lambda: o.ordinary_kriging(vario),
          "uk": lambda: o.universal_kriging(vario, deg, dim)}[kind]()
    cls = {0: gsm.GaussianVariogram, 1: gsm.SphericalVariogram, 2: gsm.ExponentialVariogram}[vk]
    fun = cls(sill=vario.sill, range=vario.range, nugget=vario.nugget)
    gm = {"sk": lambda: gsm.SimpleKriging(fun, float(z.mean())), "ok": lambda: gsm.OrdinaryKriging(fun),
          "uk": lambda: gsm.UniversalKriging(fun, deg, dim)}[kind]()
    # two referees on the same neighbour lists: the float64 reference algorithm and, for the ill-conditioned families
    # (UK with raw monomials, coincident samples, Gaussian), its binary128 evaluation
    if kind == "uk" or dups or vk == 0:
        mu, var, st, mu_q, var_q, ok_q = oc.fitpredict_exact(om, X, z, og, neighborhood=nbh, count=count, nthreads=8, **kw)
    else:
        mu, var, st, _, _ = oc.fitpredict(om, X, z, og, prob=True, neighborhood=nbh, count=count, nthreads=8, **kw)
        mu_q, var_q, ok_q = mu, var, st == 0
    try:
        got = gsm.fitpredict(gm, tab, grid, prob=True, neighborhood=nbh, count=-1 if count is None else count, **kw).z
    except ValueError as ex:
        # non-positive variance (the reference throws where Normal(mean, var) sees one).  At an exact hit the true
        # variance is 1e-10 while both solvers carry rounding noise of ~1e-15 * sill * condition: with sill = 1e13 the
        # sign of the computed value is luck on either side, so the oracle only has to be AT rounding level somewhere
        return bool(np.any((var <= 1e-9 * max(sill, 1.0)) & (st == 0))), desc + f" sill={sill} [{ex}]"
    desc += f" sill={sill:g}"
    gmu, gvar = np.ma.filled(got.mu, np.nan), np.ma.filled(got.sigma, np.nan)
    bad_o = st != 0
    bad_g = ~np.isfinite(gmu)
    both = ~bad_o & ~bad_g & ok_q
    # Numerically singular systems (coincident samples, barely determined drifts): the oracle flags exact zero
    # pivots only, the GPU everything below GSM_PIVOT_RTOL / GSM_SCHUR_RTOL -- it may flag more, within reason
    allow = 1.0 if dups else (0.15 if (kind == "uk" and nbh is not None) else 0.02)   # coincident samples: any number
    ok = np.array_equal(st == 1, np.ma.getmaskarray(got.mu) & (st == 1)) and (bad_g & ~bad_o).mean() <= allow
    # The bar, per target: within 1e-10 of the float64 reference algorithm, OR at least as close to the exact
    # (binary128) solution of the reference's equations as the float64 reference algorithm itself (ill-conditioned
    # systems -- UK with raw monomials, coincident samples -- where the reference's own rounding exceeds 1e-10).
    tol = 1e-10
    zs = max(1.0, float(np.abs(mu[both]).max())) if both.any() else 1.0
    vs = max(1.0, sill)
    def fine(g, ref, exact, scale):
        near_ref = np.abs(g - ref) <= tol * scale + tol * np.abs(ref)
        if near_ref.all():
            return True
        # the targets that miss the float64 reference: no farther from the exact solution than the float64
        # reference algorithm gets anywhere in this case (its own rounding on this family of systems)
        worst_ref = float(np.abs(ref - exact).max())
        return bool((np.abs(g - exact)[~near_ref] <= max(worst_ref, tol * scale)).all())
    ok = ok and fine(gmu[both], mu[both], mu_q[both], zs) and fine(gvar[both], var[both], var_q[both], vs)
    if not ok and both.any():
        desc += " maxerr vs f64 mu %.2e var %.2e, vs exact mu %.2e var %.2e (f64 oracle vs exact %.2e %.2e) flagged o/g %d/%d" % (
            np.abs(gmu[both] - mu[both]).max(), np.abs(gvar[both] - var[both]).max(), np.abs(gmu[both] - mu_q[both]).max(),
            np.abs(gvar[both] - var_q[both]).max(), np.abs(mu[both] - mu_q[both]).max(), np.abs(var[both] - var_q[both]).max(),
            bad_o.sum(), bad_g.sum())
    return ok, desc


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=None)
    args = ap.parse_args()
    seeds = [args.seed] if args.seed is not None else list(range(1000, 1000 + args.cases))
    t0 = time.time()
    nfail = 0
    for sd in seeds:
        try:
            ok, desc = one(sd)
        except Exception as ex:
            ok, desc = False, f"seed {sd}: EXCEPTION {ex!r}"
        if not ok:
            nfail += 1
            print("FAIL", desc, flush=True)
        elif args.seed is not None:
            print("ok  ", desc)
    print(f"{len(seeds) - nfail} of {len(seeds)} cases agree with the oracle ({time.time() - t0:.0f} s)")
