"""The oracle against every golden vector the reference's own tests hold for this path
(SURVEY.md section 8c).  CPU only."""
import numpy as np
import pytest

PSET = np.array([[25.0, 25.0], [50.0, 75.0], [75.0, 50.0]])
X0 = np.array([50.0, 50.0])


def test_docstring_simple_kriging(oracle):  # reference test/krig.jl:456-462
    o = oracle
    sph = o.Variogram(o.SPHERICAL)
    mu, _ = o.KrigingSystem(o.simple_kriging(sph, 5.0), PSET, np.array([1.0, 0.0, 0.0])).predict(X0)
    assert mu == pytest.approx(5.0, rel=1e-14)
    # I(2) * SphericalVariogram decouples into two univariate systems with means 5 and 10
    mu2, _ = o.KrigingSystem(o.simple_kriging(sph, 10.0), PSET, np.array([0.0, 1.0, 0.0])).predict(X0)
    assert mu2 == pytest.approx(10.0, rel=1e-14)


def test_docstring_ordinary_kriging(oracle):  # test/krig.jl:464-470
    o = oracle
    sph = o.Variogram(o.SPHERICAL)
    for z in ([1.0, 0.0, 0.0], [0.0, 1.0, 0.0]):
        mu, _ = o.KrigingSystem(o.ordinary_kriging(sph), PSET, np.array(z)).predict(X0)
        assert mu == pytest.approx(1 / 3, rel=1e-14)


def test_docstring_universal_kriging(oracle):  # test/krig.jl:472-478
    o = oracle
    sph = o.Variogram(o.SPHERICAL)
    uk1 = o.universal_kriging_exps(sph, [(0, 0), (1, 0), (0, 1)])
    mu, _ = o.KrigingSystem(uk1, PSET, np.array([1.0, 0.0, 0.0])).predict(X0)
    assert mu == pytest.approx(1 / 3, rel=1e-14)
    uk2 = o.universal_kriging_exps(sph, [(0, 0), (2, 0)])
    mua, _ = o.KrigingSystem(uk2, PSET, np.array([1.0, 0.0, 0.0])).predict(X0)
    mub, _ = o.KrigingSystem(uk2, PSET, np.array([0.0, 1.0, 0.0])).predict(X0)
    assert mua == pytest.approx(0.4081632653061225, rel=1e-13)
    assert mub == pytest.approx(0.34693877551020413, rel=1e-13)


@pytest.mark.parametrize("neighbors", [False, True])
def test_misc_fitpredict_spherical_grid(oracle, neighbors):  # test/misc.jl:13-25
    o = oracle
    g = o.Grid((0.5, 0.5), (1.0, 1.0), (100, 100))
    m = o.ordinary_kriging(o.Variogram(o.SPHERICAL, range=35.0))
    mu, var, st = o.fitpredict(m, PSET, np.array([1.0, 0.0, 1.0]), g, neighbors=neighbors, prob=True)
    lin = lambda i, j: (i - 1) + (j - 1) * 100  # LinearIndices, 1-based (i, j)
    assert mu[lin(25, 25)] == pytest.approx(1.0, abs=1e-3)
    assert mu[lin(50, 75)] == pytest.approx(0.0, abs=1e-3)
    assert mu[lin(75, 50)] == pytest.approx(1.0, abs=1e-3)
    assert var[lin(25, 25)] == pytest.approx(1e-10, abs=1e-14)  # exact hit: only the +1e-10 of krig.jl:278
    assert (st == 0).all()


def test_misc_idw_reproduces_data(oracle):  # test/misc.jl:2-8 (zero-distance branch idw.jl:68-69)
    o = oracle
    P = np.array([[25.0, 25.0], [50.0, 75.0], [75.0, 25.0]])
    mu, _, _ = o.fitpredict(o.IDWModel(1), P, np.array([1.0, 2.0, 3.0]), P, neighbors=False)
    assert (mu == np.array([1.0, 2.0, 3.0])).all()


def test_idw_exact_hits(oracle):  # test/idw.jl:9-16 (Euclidean analogue)
    o = oracle
    P = np.array([[1.0], [2.0], [3.0]])
    z = np.array([1.0, 0.0, 1.0])
    for j in range(3):
        assert o.idw_predict(P, z, P[j], 1) == z[j]


def test_monomial_exponent_order(oracle):  # universal.jl:14-20 + 68-77
    o = oracle
    assert o.monomial_exponents(1, 2) == [(1, 0), (0, 1), (0, 0)]
    assert o.monomial_exponents(2, 2) == [(2, 0), (0, 2), (1, 0), (0, 1), (1, 1), (0, 0)]
    assert len(o.monomial_exponents(2, 3)) == 10
    assert o.monomial_exponents(0, 3) == [(0, 0, 0)]


def test_kriging_properties(oracle):  # test/krig.jl:17-126 restated on the oracle
    o = oracle
    rng = np.random.default_rng(2024)
    X = rng.uniform(0, 10, size=(10, 3))
    z = rng.uniform(size=10)
    g = o.Variogram(o.GAUSSIAN, 1.0, 0.0, 1.0)
    models = {"sk": o.simple_kriging(g, z.mean()), "ok": o.ordinary_kriging(g), "uk": o.universal_kriging(g, 1, 3),
              "uk0": o.universal_kriging(g, 0, 3)}
    fitted = {k: o.KrigingSystem(m, X, z) for k, m in models.items()}
    for j in range(10):  # interpolation + non-negative variance + SK <= OK
        res = {k: f.predict(X[j]) for k, f in fitted.items()}
        for k, (mu, var) in res.items():
            assert mu == pytest.approx(z[j], rel=1e-6, abs=1e-6)
            assert var >= 0
        assert res["sk"][1] <= res["ok"][1] + 10 * np.finfo(float).eps
    p0 = rng.uniform(size=3) * 10
    base = {k: f.predict(p0) for k, f in fitted.items()}
    assert base["ok"][0] == pytest.approx(base["uk0"][0], rel=1e-10)   # OK == UK(deg 0)
    assert base["ok"][1] == pytest.approx(base["uk0"][1], rel=1e-10)
    h = rng.uniform(size=3)
    for k, m in models.items():  # translation invariance
        mu, var = o.KrigingSystem(m, X + h, z).predict(p0 + h)
        assert mu == pytest.approx(base[k][0], rel=1e-7)
        assert var == pytest.approx(base[k][1], rel=1e-6, abs=1e-9)
    g2 = o.Variogram(o.GAUSSIAN, 2.0, 0.0, 1.0)  # sill scaling
    mu, var = o.KrigingSystem(o.ordinary_kriging(g2), X, z).predict(p0)
    assert mu == pytest.approx(base["ok"][0], rel=1e-6)
    assert var - 1e-10 == pytest.approx(2.0 * (base["ok"][1] - 1e-10), rel=1e-5)


def test_knn_total_order_with_ties(oracle):
    o = oracle
    X = np.array([[0.0, 1.0], [1.0, 0.0], [0.0, -1.0], [-1.0, 0.0], [2.0, 0.0]])
    idx = o.knn(X, np.array([0.0, 0.0]), 3)
    assert idx.tolist() == [0, 1, 2]  # four equidistant samples: lowest rows win
    assert o.knn(X, np.array([0.0, 0.0]), 5, radius=1.5).tolist() == [0, 1, 2, 3]


def test_missing_data_reference_checks(oracle):
    """test/krig.jl:284-312 "Missing data": z = [missing, 2, 3], GaussianVariogram(), target (50, 50):
    SK / OK / DK means ~ 0 (atol 1e-6), 1 < UK(deg 1) mean < 2, all variances >= 0.  Pins the oracle's restatement of
    missingindices / lhsmissings! / rhsmissings! / the SVD solve (krig.jl:157-161, 182-209, 261-262, 381-388)."""
    o = oracle
    P = np.array([[25.0, 25.0], [50.0, 75.0], [75.0, 50.0]])
    z = np.array([np.nan, 2.0, 3.0])
    g = o.Variogram(o.GAUSSIAN, 1.0, 0.0, 1.0)
    x0 = np.array([50.0, 50.0])
    res = {}
    for name, m in (("sk", o.simple_kriging(g, 0.0)), ("ok", o.ordinary_kriging(g)), ("uk", o.universal_kriging(g, 1, 2)),
                    ("dk", o.KrigingModel(o.UK, g, 0.0, ((0, 0),)))):
        ks = o.KrigingSystem(m, P, z)
        assert ks.status()                      # _status(FHS::SVD) = true
        res[name] = ks.predict(x0)
    assert abs(res["sk"][0]) < 1e-6 and abs(res["ok"][0]) < 1e-6 and abs(res["dk"][0]) < 1e-6
    assert 1.0 < res["uk"][0] < 2.0
    assert all(v[1] >= 0 for v in res.values())
