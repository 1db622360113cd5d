"""SURVEY section 8(f) row 4, first half: CoKriging (nvar > 1) for the multivariate function the reference's tests use,
`B * gamma0` (test/krig.jl:348-395).  The oracle restates the reference's FULL interleaved system
((nobs nvar + ncon nvar) unknowns, oracle CoKrigingSystem); the CUDA path uses that the proportional model with
isotopic data decouples into one univariate solve per variable (autokrigeability) -- the two must agree."""
import numpy as np
import pytest

B3 = np.array([[1.0, 0.3, 0.1], [0.3, 1.0, 0.2], [0.1, 0.2, 1.0]])


def test_cokriging_oracle_reference_checks(oracle):
    """test/krig.jl:348-395 on the oracle: interpolation property for SK / OK / UK(1, 2) / DK, variances ~ 0 at the
    samples, means in [0, 1] and variances >= 0 at (50, 50), SK with non-zero means."""
    o = oracle
    P = np.array([[25.0, 25.0], [50.0, 75.0], [75.0, 50.0]])
    Z = np.eye(3)
    g = o.Variogram(o.SPHERICAL, 1.0, 0.0, 35.0)
    for kind, kw in ((o.SK, dict(mean=[0.0, 0.0, 0.0])), (o.OK, {}), (o.UK, dict(exps=o.monomial_exponents(1, 2))),
                     (o.UK, dict(exps=((0, 0),))), (o.SK, dict(mean=[0.7, 0.2, 0.1]))):
        ks = o.CoKrigingSystem(kind, B3, g, P, Z, **kw)
        for i in range(3):
            mu, var = ks.predict(P[i])
            assert np.allclose(mu, np.eye(3)[i], atol=1e-9) and np.allclose(var, 0.0, atol=1e-8)
        mu, var = ks.predict(np.array([50.0, 50.0]))
        assert (var >= 0).all()
        if kw.get("mean", [0.0])[0] == 0.0:
            assert ((mu >= -1e-12) & (mu <= 1 + 1e-12)).all()


def test_proportional_cokriging_decouples_on_the_oracle(oracle):
    """the property the GPU path relies on, on the CPU: full cokriging system == univariate systems with B_pp gamma0"""
    o = oracle
    rng = np.random.default_rng(5)
    X = rng.uniform(0, 30, size=(14, 2))
    Z = rng.standard_normal((14, 2))
    B = np.array([[2.0, 0.7], [0.7, 0.5]])
    g = o.Variogram(o.EXPONENTIAL, 1.3, 0.2, 9.0)
    for kind, kw in ((o.SK, dict(mean=[0.3, -0.2])), (o.OK, {}), (o.UK, dict(exps=o.monomial_exponents(1, 2)))):
        ks = o.CoKrigingSystem(kind, B, g, X, Z, **kw)
        for x0 in rng.uniform(0, 30, size=(5, 2)):
            mu, var = ks.predict(x0)
            for p_ in range(2):
                v1 = o.Variogram(o.EXPONENTIAL, 1.3 * B[p_, p_], 0.2 * B[p_, p_], 9.0)
                m1 = {o.SK: lambda: o.simple_kriging(v1, kw.get("mean", [0, 0])[p_]), o.OK: lambda: o.ordinary_kriging(v1),
                      o.UK: lambda: o.universal_kriging(v1, 1, 2)}[kind]()
                mu1, var1 = o.KrigingSystem(m1, X, Z[:, p_]).predict(x0)
                assert abs(mu1 - mu[p_]) < 1e-10 and abs(var1 - var[p_]) < 1e-10


def test_mvnormal_host_class(gsm):
    d = gsm.MvNormal([1.0, 2.0], [[2.0, 0.3], [0.3, 1.0]])
    assert list(d.mean()) == [1.0, 2.0] and list(d.var()) == [2.0, 1.0] and d.cov().shape == (2, 2)
    with pytest.raises(np.linalg.LinAlgError):       # PosDefException of MvNormal(mu, Sigma), krig.jl:236
        gsm.MvNormal([0.0, 0.0], [[1.0, 2.0], [2.0, 1.0]])


@pytest.mark.gpu
def test_cokriging_gpu_matches_full_system(gsm, oracle):
    o = oracle
    rng = np.random.default_rng(8)
    n = 400
    X = rng.uniform(0, 40, size=(n, 2))
    Z = np.stack([np.sin(X[:, 0] / 7), np.cos(X[:, 1] / 5), 0.3 * X[:, 0] / 40], axis=1) + 0.05 * rng.standard_normal((n, 3))
    base_g, base_o = gsm.SphericalVariogram(range=12.0, nugget=0.1), o.Variogram(o.SPHERICAL, 1.0, 0.1, 12.0)
    fun = gsm.CoregionalizedVariogram(B3, base_g)
    tab = gsm.georef({"a": Z[:, 0], "b": Z[:, 1], "c": Z[:, 2]}, gsm.PointSet(X))
    P = rng.uniform(2, 38, size=(60, 2))
    k = 12
    for gm, kind, kw in ((gsm.OrdinaryKriging(fun), o.OK, {}), (gsm.SimpleKriging(fun, [0.1, 0.2, 0.3]), o.SK, dict(mean=[0.1, 0.2, 0.3])),
                         (gsm.UniversalKriging(fun, 1, 2), o.UK, dict(exps=o.monomial_exponents(1, 2)))):
        pred = gsm.fitpredict(gm, tab, gsm.PointSet(P), maxneighbors=k, prob=True)
        alpha = o.scale_factor(12.0, o.points_absmax(X), o.points_absmax(P))
        sX, sP, sv = X * alpha, P * alpha, base_o.scaled(alpha)
        for j in range(len(P)):
            idx = o.knn(sX, sP[j], k)
            mu, var = o.CoKrigingSystem(kind, B3, sv, sX[idx], Z[idx], **kw).predict(sP[j])
            for p_, name in enumerate(("a", "b", "c")):
                col = pred.table[name]
                assert abs(col.mu[j] - mu[p_]) <= 1e-9 * max(1.0, abs(mu[p_]))
                assert abs(col.sigma[j] - var[p_]) <= 1e-9 * max(1.0, abs(var[p_]))
    with pytest.raises(ValueError):
        gsm.fitpredict(gsm.OrdinaryKriging(fun), gsm.georef({"a": Z[:, 0]}, gsm.PointSet(X)), gsm.PointSet(P))


@pytest.mark.gpu
def test_cokriging_fit_predictprob_reference_testset(gsm, oracle):
    """test/krig.jl:348-395 through fit / predict / predictprob on the CUDA path (MvNormal over the three variables),
    with the full covariance matrix checked against the reference's interleaved system restated by the oracle."""
    o = oracle
    P = np.array([[25.0, 25.0], [50.0, 75.0], [75.0, 50.0]])
    Z = np.eye(3)
    tab = gsm.georef({"a": Z[:, 0], "b": Z[:, 1], "c": Z[:, 2]}, gsm.PointSet(P))
    fun = gsm.CoregionalizedVariogram(B3, gsm.SphericalVariogram(range=35.0))
    g = o.Variogram(o.SPHERICAL, 1.0, 0.0, 35.0)
    vars_ = ("a", "b", "c")
    for gm, kind, kw in ((gsm.SimpleKriging(fun, [0.0, 0.0, 0.0]), o.SK, dict(mean=[0.0, 0.0, 0.0])),
                         (gsm.OrdinaryKriging(fun), o.OK, {}),
                         (gsm.UniversalKriging(fun, 1, 2), o.UK, dict(exps=o.monomial_exponents(1, 2))),
                         (gsm.UniversalKriging(fun, exps=((0, 0),)), o.UK, dict(exps=((0, 0),)))):
        f = gsm.fit(gm, tab)
        assert gsm.status(f)
        ks = o.CoKrigingSystem(kind, B3, g, P, Z, **kw)
        for i in range(3):                                  # interpolation property
            d = gsm.predictprob(f, vars_, P[i])
            assert np.allclose(d.mean(), np.eye(3)[i], atol=1e-9) and np.allclose(d.var(), 0.0, atol=1e-8)
        d = gsm.predictprob(f, vars_, (50.0, 50.0))
        assert ((d.mean() >= 0) & (d.mean() <= 1)).all() and (d.var() >= 0).all()
        mu, Sigma = ks.predict_full(np.array([50.0, 50.0]))
        assert np.allclose(d.mean(), mu, rtol=1e-10, atol=1e-12)
        assert np.allclose(d.cov(), Sigma, rtol=1e-9, atol=1e-12), np.abs(d.cov() - Sigma).max()
        assert np.allclose(gsm.predict(f, vars_, (50.0, 50.0)), mu, rtol=1e-10, atol=1e-12)
    f = gsm.fit(gsm.SimpleKriging(fun, [0.7, 0.2, 0.1]), tab)   # simple cokriging with non-zero means
    for i in range(3):
        assert np.allclose(gsm.predict(f, vars_, P[i]), np.eye(3)[i], atol=1e-9)
    with pytest.raises(ValueError):
        gsm.predict(f, ("a", "b"), P[0])
    # the columns a `vars` tuple binds to the variables may be permuted (krigmean indexes vars[p], krig.jl:250-256)
    ks = o.CoKrigingSystem(o.SK, B3, g, P, Z[:, [1, 0, 2]], mean=[0.7, 0.2, 0.1])
    mu, _ = ks.predict_full(np.array([40.0, 60.0]))
    assert np.allclose(gsm.predict(f, ("b", "a", "c"), (40.0, 60.0)), mu, rtol=1e-10, atol=1e-12)
