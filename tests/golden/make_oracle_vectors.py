#!/usr/bin/env python
"""Generates tests/golden/oracle_vectors.npz: seeded inputs and the ORACLE's outputs for a set of small cases.
The reference is a Julia package and no Julia exists in this image, so these vectors pin the oracle against itself
(regression) and give the CUDA path committed fixtures to match; the reference's own golden numbers are in
reference_goldens.json.  Re-run from the repo root:  python tests/golden/make_oracle_vectors.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
o = entry.load_oracle()

CASES = [
    # name, kind, variogram kind, dim, degree, k (0 = global), grid dims, n samples
    ("ok_sph_2d_k16", "ok", o.SPHERICAL, 2, None, 16, (12, 9), 400),
    ("sk_exp_2d_k8", "sk", o.EXPONENTIAL, 2, None, 8, (10, 7), 300),
    ("uk1_gau_2d_k24", "uk", o.GAUSSIAN, 2, 1, 24, (9, 8), 500),
    ("ok_sph_3d_k32", "ok", o.SPHERICAL, 3, None, 32, (6, 5, 4), 900),
    ("uk2_sph_3d_k40", "uk", o.SPHERICAL, 3, 2, 40, (5, 4, 3), 900),
    ("ok_gau_2d_global", "ok", o.GAUSSIAN, 2, None, 0, (8, 6), 150),
    ("idw2_2d_k12", "idw", None, 2, None, 12, (11, 6), 350),
    ("idw1_3d_global", "idw1", None, 3, None, 0, (5, 5, 4), 200),
]


def build(case):
    name, kind, vk, dim, deg, k, dims, n = case
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    X = rng.uniform(0, 50, size=(n, dim))
    z = np.sin(X[:, 0] / 9) + np.cos(X[:, -1] / 13) + 0.1 * rng.standard_normal(n)
    grid = o.Grid(tuple(0.0 for _ in dims), tuple(50.0 / d for d in dims), dims)
    if kind.startswith("idw"):
        model = o.IDWModel(2.0 if kind == "idw" else 1.0)
    else:
        vario = o.Variogram(vk, 1.0, 0.1, 18.0)
        model = {"sk": lambda: o.simple_kriging(vario, float(z.mean())), "ok": lambda: o.ordinary_kriging(vario),
                 "uk": lambda: o.universal_kriging(vario, deg, dim)}[kind]()
    prob = not kind.startswith("idw")
    if k:
        mu, var, st = o.fitpredict(model, X, z, grid, prob=prob, maxneighbors=k)
    else:
        mu, var, st = o.fitpredict(model, X, z, grid, prob=prob, neighbors=False)
    out = {"X": X, "z": z, "mean": mu, "var": var if var is not None else np.zeros(0), "status": st}
    if k and prob:
        # the exact (binary128, oracle/gsm_oracle_q.c) solution of the same equations on the same neighbour lists:
        # the referee for the ill-conditioned UniversalKriging cases, where float64 Bunch-Kaufman on raw monomials
        # is itself several digits away from the solution of its own system
        import gsm_oracle_c as oc
        _, _, _, mu_q, var_q, ok_q = oc.fitpredict_exact(model, X, z, grid, maxneighbors=k, nthreads=4)
        assert ok_q.all()
        out["mean_exact"], out["var_exact"] = mu_q, var_q
    return out


if __name__ == "__main__":
    out = {}
    for case in CASES:
        for key, val in build(case).items():
            out[f"{case[0]}/{key}"] = val
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
