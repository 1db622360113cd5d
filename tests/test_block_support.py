"""SURVEY section 8(f) row 4, second half: change of support (`fitpredict(...; point=false)`, models.jl:114-115,136,205;
`predictprob(fitted, :z, Quadrangle)`, test/krig.jl:221-242).  The right-hand side becomes the mean covariance between
a sample and the points that stand for the target, covzero the mean covariance among those points (krig.jl:295-301,
342-378); the search centre and the drift functions stay at the centroid (models.jl:164, universal.jl:125-126).

Which points stand for a cell is GeoStatsFunctions' rule, not in the reference tree: the arithmetic GIVEN the points
is checked against the numpy oracle at 1e-10, the rule itself is restated twice (product / oracle) and compared; the
only thing the reference's own tests pin is var >= 0, restated below on the oracle (CPU) and on the CUDA path."""
import numpy as np
import pytest

RTOL, ATOL = 1e-10, 1e-12


# ---- CPU: the oracle ------------------------------------------------------------------------------
def _support_data():
    return np.array([[25.0, 25.0], [50.0, 75.0], [75.0, 50.0]]), np.array([1.0, 0.0, 0.0])


def test_oracle_change_of_support_variances(oracle):
    """test/krig.jl:221-242: Gaussian(sill 1, range 1, nugget 0), unit quadrangle, SK / OK / UK(1, 2) / DK([1])."""
    o = oracle
    X, z = _support_data()
    v = o.Variogram(o.GAUSSIAN, 1.0, 0.0, 1.0)
    blk = o.block_offsets([1.0, 1.0], 1.0)
    assert blk.shape == (9, 2)
    c = np.array([0.5, 0.5])
    for model in (o.simple_kriging(v, z.mean()), o.ordinary_kriging(v), o.universal_kriging(v, 1, 2),
                  o.universal_kriging_exps(v, [(0, 0)])):
        mu, var = o.KrigingSystem(model, X, z).predict(c, blk)
        assert np.isfinite(mu) and var >= 0


def test_oracle_single_point_block_is_point_support(oracle):
    o = oracle
    rng = np.random.default_rng(3)
    X = rng.uniform(0, 10, (12, 2))
    z = rng.standard_normal(12)
    ks = o.KrigingSystem(o.ordinary_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.1, 4.0)), X, z)
    x0 = np.array([4.2, 5.1])
    assert ks.predict(x0) == ks.predict(x0, np.zeros((1, 2)))


def test_oracle_block_mean_is_mean_of_point_means(oracle):
    """One system, right-hand sides averaged: the block estimate is the average of the point estimates at the block's
    points whenever the constraint rows agree (SK, OK; linear drift on a lattice symmetric about the centroid)."""
    o = oracle
    rng = np.random.default_rng(4)
    X = rng.uniform(0, 10, (15, 2))
    z = rng.standard_normal(15)
    v = o.Variogram(o.EXPONENTIAL, 1.2, 0.05, 5.0)
    blk = o.block_offsets([1.5, 0.8], 5.0)
    x0 = np.array([5.0, 4.0])
    for model in (o.simple_kriging(v, 0.3), o.ordinary_kriging(v), o.universal_kriging(v, 1, 2)):
        ks = o.KrigingSystem(model, X, z)
        mu_b, var_b = ks.predict(x0, blk)
        mus = [ks.predict(x0 + q)[0] for q in blk]
        assert abs(mu_b - np.mean(mus)) < 1e-12
        assert var_b <= ks.predict(x0)[1] + 1e-12       # averaging over the block cannot add variance here


def test_block_rule_restated_twice(gsm, oracle):
    for sides, r in (([1.0, 1.0], 1.0), ([1.0, 2.0], 10.0), ([1.0, 1.0, 1.0], 0.5), ([0.5], 3.0), ([2.0, 1.0, 0.7], 1.0)):
        a = oracle.block_offsets(sides, r)
        b = gsm.block_points(gsm.Box([0.0] * len(sides), sides), r)
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-15
    g = gsm.CartesianGrid((0.0, 0.0), (8.0, 4.0), dims=(4, 8))
    assert np.abs(gsm.block_points(g, 100.0) - oracle.block_offsets([2.0, 0.5], 100.0)).max() < 1e-15
    q = gsm.Quadrangle((0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0))
    assert q.sides == (1.0, 1.0) and tuple(q.centroid()) == (0.5, 0.5)
    with pytest.raises(NotImplementedError):
        gsm.Quadrangle((0.0, 0.0), (1.0, 0.2), (1.0, 1.0), (0.0, 1.0))


# ---- GPU ---------------------------------------------------------------------------------------------
def _problem(dim, n, seed):
    rng = np.random.default_rng(seed)
    ext = np.array([40.0, 30.0, 20.0])[:dim]
    X = rng.uniform(0, 1, size=(n, dim)) * ext
    z = np.sin(X[:, 0] / 6) + np.cos(X[:, -1] / 5) + 0.1 * rng.standard_normal(n)
    return X, z, ext


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 3])
def test_local_block_kriging_against_oracle(gsm, oracle, dim):
    o = oracle
    X, z, ext = _problem(dim, 1500, 10 + dim)
    dims = (13, 11, 5)[:dim]
    grid = gsm.CartesianGrid((0.0,) * dim, tuple(ext), dims=dims)
    og = o.Grid((0.0,) * dim, tuple(e / d for e, d in zip(ext, dims)), dims)
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    nested = 0.6 * gsm.SphericalVariogram(range=5.0, nugget=0.05) + 0.4 * gsm.GaussianVariogram(range=15.0)
    onested = o.Variogram(o.SPHERICAL, structs=((o.SPHERICAL, 0.6, 0.03, 5.0, None), (o.GAUSSIAN, 0.4, 0.0, 15.0, None)))
    cases = (
        (gsm.SphericalVariogram(range=9.0, sill=1.3, nugget=0.1), o.Variogram(o.SPHERICAL, 1.3, 0.1, 9.0)),
        (gsm.GaussianVariogram(range=2.0, sill=1.0, nugget=0.02), o.Variogram(o.GAUSSIAN, 1.0, 0.02, 2.0)),
        (nested, onested),
    )
    for gv, ov in cases:
        blk = gsm.block_points(grid, gv.range)
        for gm, om, kw in ((gsm.SimpleKriging(gv, 0.2), o.simple_kriging(ov, 0.2), dict(maxneighbors=12)),
                           (gsm.OrdinaryKriging(gv), o.ordinary_kriging(ov), dict(maxneighbors=32)),
                           (gsm.OrdinaryKriging(gv), o.ordinary_kriging(ov), dict(maxneighbors=44)),
                           (gsm.UniversalKriging(gv, 1, dim), o.universal_kriging(ov, 1, dim), dict(maxneighbors=20, minneighbors=3,
                                                                                                  neighborhood=9.0))):
            mu, var, st = o.fitpredict(om, X, z, og, prob=True, block_offsets=blk, **kw)
            got = gsm.fitpredict(gm, tab, grid, point=False, prob=True, **kw).z
            ok = st == 0
            assert ok.sum() > 0.5 * len(st)
            gmu, gvar = np.ma.filled(got.mu, np.nan), np.ma.filled(got.sigma, np.nan)
            assert np.array_equal(np.isnan(gmu), ~ok)
            assert np.allclose(gmu[ok], mu[ok], rtol=RTOL, atol=ATOL), np.abs(gmu[ok] - mu[ok]).max()
            assert np.allclose(gvar[ok], var[ok], rtol=RTOL, atol=ATOL), np.abs(gvar[ok] - var[ok]).max()
            # and it is not the point-support answer
            pt = gsm.fitpredict(gm, tab, grid, point=True, prob=True, **kw).z
            assert np.nanmax(np.abs(np.ma.filled(pt.sigma, np.nan)[ok] - gvar[ok])) > 1e-6


@pytest.mark.gpu
def test_global_block_kriging_against_oracle_and_linearity(gsm, oracle):
    o = oracle
    X, z, ext = _problem(2, 400, 21)
    dims = (17, 9)
    grid = gsm.CartesianGrid((0.0, 0.0), tuple(ext), dims=dims)
    og = o.Grid((0.0, 0.0), tuple(e / d for e, d in zip(ext, dims)), dims)
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    gv, ov = gsm.ExponentialVariogram(range=12.0, sill=0.9, nugget=0.05), o.Variogram(o.EXPONENTIAL, 0.9, 0.05, 12.0)
    blk = gsm.block_points(grid, gv.range)
    for gm, om in ((gsm.SimpleKriging(gv, 0.1), o.simple_kriging(ov, 0.1)), (gsm.OrdinaryKriging(gv), o.ordinary_kriging(ov)),
                   (gsm.UniversalKriging(gv, 2, 2), o.universal_kriging(ov, 2, 2))):
        mu, var, st = o.fitpredict(om, X, z, og, prob=True, neighbors=False, block_offsets=blk)
        got = gsm.fitpredict(gm, tab, grid, point=False, prob=True, neighbors=False).z
        assert (st == 0).all()
        assert np.allclose(got.mu, mu, rtol=1e-8, atol=1e-10), np.abs(got.mu - mu).max()
        assert np.allclose(got.sigma, var, rtol=1e-8, atol=1e-10), np.abs(got.sigma - var).max()
    # linearity at a size the oracle does not reach: the block estimate is the average of the point estimates at the
    # block's points (one global system, averaged right-hand sides; SK and OK)
    X, z, ext = _problem(2, 3000, 22)
    dims = (64, 48)
    grid = gsm.CartesianGrid((0.0, 0.0), tuple(ext), dims=dims)
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    blk = gsm.block_points(grid, gv.range)
    cent = (np.stack(np.meshgrid(*[np.arange(d) + 0.5 for d in dims], indexing="ij"), -1).reshape(-1, 2, order="F")
            * np.array(grid.spacing))
    for gm in (gsm.SimpleKriging(gv, 0.1), gsm.OrdinaryKriging(gv)):
        b = gsm.fitpredict(gm, tab, grid, point=False, neighbors=False).z
        acc = np.zeros(len(cent))
        for q in blk:
            # (same scale factor as the grid call: the bounding box of the shifted points stays inside the grid's)
            acc += gsm.fitpredict(gm, tab, gsm.PointSet(np.clip(cent + q, 0.0, ext)), neighbors=False).z
        assert np.allclose(b, acc / len(blk), rtol=1e-9, atol=1e-10), np.abs(b - acc / len(blk)).max()


@pytest.mark.gpu
def test_predictprob_on_a_quadrangle(gsm, oracle):
    """test/krig.jl:221-242 on the CUDA path, with the values checked against the oracle beside the var >= 0 of the
    reference's test."""
    o = oracle
    X, z = _support_data()
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    gv, ov = gsm.GaussianVariogram(sill=1.0, range=1.0, nugget=0.0), o.Variogram(o.GAUSSIAN, 1.0, 0.0, 1.0)
    quad = gsm.Quadrangle((0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0))
    blk = o.block_offsets([1.0, 1.0], 1.0)
    for gm, om in ((gsm.SimpleKriging(gv, float(z.mean())), o.simple_kriging(ov, z.mean())),
                   (gsm.OrdinaryKriging(gv), o.ordinary_kriging(ov)),
                   (gsm.UniversalKriging(gv, 1, 2), o.universal_kriging(ov, 1, 2)),
                   (gsm.UniversalKriging(gv, exps=((0, 0),)), o.universal_kriging_exps(ov, [(0, 0)]))):
        f = gsm.fit(gm, tab)
        d = gsm.predictprob(f, "z", quad)
        mu, var = o.KrigingSystem(om, X, z).predict(np.array([0.5, 0.5]), blk)
        assert d.sigma ** 2 >= 0
        assert abs(d.mu - mu) <= 1e-10 * max(1.0, abs(mu)) and abs(d.sigma ** 2 - var) <= 1e-10
        assert abs(gsm.predict(f, "z", quad) - mu) <= 1e-10 * max(1.0, abs(mu))
        # the context is back on point support afterwards
        p = gsm.predictprob(f, "z", (50.0, 50.0))
        mu_p, var_p = o.KrigingSystem(om, X, z).predict(np.array([50.0, 50.0]))
        assert abs(p.mu - mu_p) <= 1e-10 * max(1.0, abs(mu_p)) and abs(p.sigma ** 2 - var_p) <= 1e-10


@pytest.mark.gpu
def test_block_support_boundaries(gsm, oracle):
    X, z, ext = _problem(2, 500, 31)
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    grid = gsm.CartesianGrid((0.0, 0.0), tuple(ext), dims=(16, 12))
    gv = gsm.SphericalVariogram(range=9.0, nugget=0.1)
    # IDW / NN only see centroids (idw.jl:82, nn.jl): point=false changes nothing
    for m in (gsm.IDW(2), gsm.NN()):
        a = gsm.fitpredict(m, tab, grid, point=False, maxneighbors=8).z
        b = gsm.fitpredict(m, tab, grid, point=True, maxneighbors=8).z
        assert np.array_equal(a, b)
    # the elements of a PointSet are points
    P = gsm.PointSet(np.random.default_rng(5).uniform(0, 30, (50, 2)))
    a = gsm.fitpredict(gsm.OrdinaryKriging(gv), tab, P, point=False, maxneighbors=8).z
    b = gsm.fitpredict(gsm.OrdinaryKriging(gv), tab, P, point=True, maxneighbors=8).z
    assert np.array_equal(a, b)
    # a block of one point at the centroid is point support (same kernel; the two instances of the covariance function
    # may contract their multiply-adds differently, hence not bit for bit)
    ctx = gsm.default_context()
    ctx.set_option("local_impl", "simt")
    try:
        a = gsm.fitpredict(gsm.OrdinaryKriging(gv), tab, grid, point=False, prob=True, maxneighbors=16,
                           block_offsets=np.zeros((1, 2))).z
        b = gsm.fitpredict(gsm.OrdinaryKriging(gv), tab, grid, point=True, prob=True, maxneighbors=16).z
    finally:
        ctx.set_option("local_impl", "auto")
    assert np.allclose(a.mu, b.mu, rtol=1e-12, atol=1e-13) and np.allclose(a.sigma, b.sigma, rtol=1e-12, atol=1e-13)
    # too many points per cell is refused, not truncated
    with pytest.raises(NotImplementedError):
        gsm.fitpredict(gsm.OrdinaryKriging(gsm.SphericalVariogram(range=0.05)), tab, grid, point=False, maxneighbors=8)
    with pytest.raises(gsm.GsmError):
        ctx.set_block_support(np.full((2, 2), np.nan))
    ctx.set_block_support(None)
