"""The binary128 twin of the oracle (oracle/gsm_oracle_q.c) is the exact-arithmetic referee of the ill-conditioned
configs (UniversalKriging with raw monomials, C4).  Pin it: (1) on well-conditioned systems it agrees with the
float64 oracle to rounding, (2) it reproduces the reference's golden values, (3) it agrees with 50-digit mpmath
arithmetic on C4-class systems to far below any float64 resolution."""
import numpy as np
import pytest


def _oc():
    import gsm_oracle_c as oc
    return oc


def test_quad_twin_matches_float64_oracle_when_well_conditioned(oracle):
    o, oc = oracle, _oc()
    rng = np.random.default_rng(21)
    X = rng.uniform(0, 100, size=(3000, 2))
    z = np.sin(X[:, 0] / 9) + 0.1 * rng.standard_normal(3000)
    grid = o.Grid((0.0, 0.0), (2.5, 2.5), (40, 40))
    for model in (o.ordinary_kriging(o.Variogram(o.SPHERICAL, 1.3, 0.1, 30.0)),
                  o.simple_kriging(o.Variogram(o.EXPONENTIAL, 0.8, 0.05, 25.0), float(z.mean())),
                  o.ordinary_kriging(o.Variogram(o.GAUSSIAN, 1.0, 0.2, 15.0))):
        idx = np.arange(0, 1600, 7)
        r = oc.grid_subset_exact(model, X, z, grid, idx, 16)
        assert r["ok_q"].all()
        assert np.allclose(r["mu64"], r["mu_q"], rtol=1e-11, atol=1e-13)
        assert np.allclose(r["var64"], r["var_q"], rtol=1e-11, atol=1e-13)


def test_quad_twin_reference_goldens(oracle):
    """test/krig.jl:451-479: SK(Spherical, mean 5) -> 5.0, OK -> 1/3, UK [1, x^2] style -> 0.40816.../0.34693..."""
    o, oc = oracle, _oc()
    P = np.array([[25.0, 25.0], [50.0, 75.0], [75.0, 50.0]])
    T = np.array([[50.0, 50.0]])
    nbr = np.array([[0, 1, 2]], dtype=np.int32)
    cnt = np.array([3], dtype=np.int32)
    sph = o.Variogram(o.SPHERICAL, 1.0, 0.0, 1.0)
    mu, _, ok, _ = oc.local_quad(o.simple_kriging(sph, 5.0), P, np.array([1.0, 0.0, 0.0]), T, nbr, cnt)
    assert ok[0] and mu[0] == pytest.approx(5.0, abs=1e-15)
    mu, _, ok, _ = oc.local_quad(o.ordinary_kriging(sph), P, np.array([1.0, 0.0, 0.0]), T, nbr, cnt)
    assert ok[0] and mu[0] == pytest.approx(1.0 / 3.0, abs=1e-15)
    uk = o.KrigingModel(o.UK, sph, 0.0, ((0, 0), (2, 0)))     # drifts [1, x^2]
    mu, _, ok, _ = oc.local_quad(uk, P, np.array([1.0, 0.0, 0.0]), T, nbr, cnt)
    assert ok[0] and mu[0] == pytest.approx(0.4081632653061225, abs=1e-15)
    mu, _, ok, _ = oc.local_quad(uk, P, np.array([0.0, 1.0, 0.0]), T, nbr, cnt)
    assert ok[0] and mu[0] == pytest.approx(0.34693877551020413, abs=1e-15)


def test_quad_twin_against_mpmath_on_c4_class_systems(oracle):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 60
    o, oc = oracle, _oc()
    rng = np.random.default_rng(4)
    n = 20000
    X = rng.uniform(0, 64, size=(n, 3))
    z = np.sin(X[:, 0] / 20) + np.cos(X[:, 2] / 30) + 0.1 * rng.standard_normal(n)
    vario = o.Variogram(o.SPHERICAL, 1.0, 0.1, 20.0)
    model = o.universal_kriging(vario, 2, 3)
    grid = o.Grid((0.0,) * 3, (1.0,) * 3, (64, 64, 64))
    idx = np.array([1234, 77777, 200001])
    r = oc.grid_subset_exact(model, X, z, grid, idx, 32)
    alpha = o.scale_factor(vario.range, o.points_absmax(X), grid.bbox_absmax())
    sv = vario.scaled(alpha)
    sX = X * alpha
    sg = grid.scaled(alpha)
    exps = model.exps

    def cov(h):
        if h == 0:
            return mp.mpf(sv.sill)
        t = h / mp.mpf(sv.range)
        g = (mp.mpf(sv.sill) - mp.mpf(sv.nugget)) * (mp.mpf("1.5") * t - mp.mpf("0.5") * t ** 3) if h < sv.range \
            else mp.mpf(sv.sill) - mp.mpf(sv.nugget)
        return mp.mpf(sv.sill) - (g + mp.mpf(sv.nugget))

    for j, lin in enumerate(idx):
        x0 = sg.centroids(int(lin), 1)[0]
        ids = r["nbr"][j]
        Q = [[mp.mpf(float(c)) for c in sX[i]] for i in ids]
        t0 = [mp.mpf(float(c)) for c in x0]
        k, p = 32, len(exps)
        A = mp.matrix(k + p, k + p)
        b = mp.matrix(k + p, 1)
        for a in range(k):
            for c in range(k):
                A[a, c] = cov(mp.sqrt(sum((Q[a][d] - Q[c][d]) ** 2 for d in range(3))))
            b[a] = cov(mp.sqrt(sum((Q[a][d] - t0[d]) ** 2 for d in range(3))))
            for q, e in enumerate(exps):
                f = mp.mpf(1)
                for d in range(3):
                    f *= Q[a][d] ** e[d]
                A[a, k + q] = f
                A[k + q, a] = f
        for q, e in enumerate(exps):
            f = mp.mpf(1)
            for d in range(3):
                f *= t0[d] ** e[d]
            b[k + q] = f
        w = mp.lu_solve(A, b)
        mu_e = sum(w[a] * mp.mpf(float(z[ids[a]])) for a in range(k))
        var_e = cov(0) - sum(b[a] * w[a] for a in range(k + p)) + mp.mpf("1e-10")
        assert abs(mp.mpf(float(r["mu_q"][j])) - mu_e) <= mp.mpf("4e-16") * max(1, abs(mu_e))
        assert abs(mp.mpf(float(r["var_q"][j])) - var_e) <= mp.mpf("4e-16") * max(1, abs(var_e))
        # ...while the float64 reference algorithm itself is several digits away from its own equations here
        assert r["st64"][j] == 0
