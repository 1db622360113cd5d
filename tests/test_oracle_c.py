"""The C twin of the oracle (used for large parity cases and as the timed CPU baseline) against the
numpy oracle, which is pinned to the reference's golden values.  CPU only."""
import os
import sys

import numpy as np
import pytest


@pytest.fixture(scope="module")
def oc(oracle):
    import gsm_oracle_c
    return gsm_oracle_c


def _data(seed, n, dim):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 100, size=(n, dim))
    z = np.sin(X[:, 0] / 9) + 0.2 * rng.standard_normal(n)
    return X, z


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_kdtree_equals_bruteforce_equals_numpy(oracle, oc, dim):
    X, z = _data(1 + dim, 3000, dim)
    T = np.random.default_rng(7).uniform(-10, 110, size=(150, dim))
    _, _, _, nbr_tree, _ = oc.local(None, X, z, T, 24, knn_only=True, use_tree=True)
    _, _, _, nbr_brute, _ = oc.local(None, X, z, T, 24, knn_only=True, use_tree=False, nthreads=3)
    assert np.array_equal(nbr_tree, nbr_brute)
    for j in range(len(T)):
        assert np.array_equal(nbr_tree[j], oracle.knn(X, T[j], 24))


def test_kdtree_ties_on_lattice(oracle, oc):
    g = np.stack(np.meshgrid(np.arange(16.0), np.arange(16.0), indexing="ij"), -1).reshape(-1, 2)
    X = g[np.random.default_rng(3).permutation(len(g))]
    T = np.concatenate([g[::5], g[::9] + 0.5])
    _, _, _, nbr, _ = oc.local(None, X, np.zeros(len(X)), T, 12, knn_only=True)
    for j in range(len(T)):
        assert np.array_equal(nbr[j], oracle.knn(X, T[j], 12))


CASES = [("sk", 1, 2, None), ("ok", 1, 2, None), ("ok", 0, 2, None), ("ok", 2, 3, None), ("uk", 1, 2, 1),
         ("uk", 1, 2, 2), ("uk", 1, 3, 2)]


@pytest.mark.parametrize("kind,vk,dim,deg", CASES)
def test_c_bunch_kaufman_matches_lapack_path(oracle, oc, kind, vk, dim, deg):
    o = oracle
    X, z = _data(40 + vk + dim, 1500, dim)
    vario = o.Variogram(vk, 1.0, 0.1, 25.0)
    model = {"sk": o.simple_kriging(vario, float(z.mean())), "ok": o.ordinary_kriging(vario)}.get(kind) \
        or o.universal_kriging(vario, deg, dim)
    dims = (12, 10) if dim == 2 else (5, 5, 4)
    grid = o.Grid(tuple(0.0 for _ in dims), tuple(100.0 / d for d in dims), dims)
    mu, var, st = o.fitpredict(model, X, z, grid, maxneighbors=20, prob=True)
    cmu, cvar, cst, _, _ = oc.fitpredict(model, X, z, grid, maxneighbors=20, prob=True, nthreads=2)
    assert (st == 0).all() and (cst == 0).all()
    assert np.allclose(cmu, mu, rtol=1e-11, atol=1e-13)
    assert np.allclose(cvar, var, rtol=1e-11, atol=1e-13)


def test_c_idw_and_missing(oracle, oc):
    o = oracle
    X, z = _data(9, 800, 2)
    P = np.concatenate([X[:10], np.random.default_rng(1).uniform(0, 100, size=(60, 2)), [[900.0, 900.0]]])
    mu, _, st = o.fitpredict(o.IDWModel(2), X, z, P, maxneighbors=8, minneighbors=2, neighborhood=15.0)
    cmu, _, cst, _, _ = oc.fitpredict(o.IDWModel(2), X, z, P, maxneighbors=8, minneighbors=2, neighborhood=15.0)
    assert np.array_equal(st, cst) and st[-1] == 1
    ok = st == 0
    assert np.allclose(cmu[ok], mu[ok], rtol=1e-13)
    assert np.array_equal(cmu[:10], z[:10])
