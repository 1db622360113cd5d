"""Missing values: reference missingindices / lhsmissings! / rhsmissings! / SVD minimum-norm solve
(krig.jl:157-161, 182-209, 261-262, 381-388).  The oracle restates exactly that (zeroed rows and columns + numpy SVD with
Julia's truncation rule); the CUDA path solves the block-eliminated form on the samples that have a value, with the
constraints projected onto null(F_m)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cmp(got, mu, var, st, rtol=1e-9):
    gmu = np.ma.filled(got.mu, np.nan) if np.ma.isMaskedArray(got.mu) else np.asarray(got.mu)
    gvar = np.ma.filled(got.sigma, np.nan) if np.ma.isMaskedArray(got.sigma) else np.asarray(got.sigma)
    ok = st == 0
    assert np.isfinite(gmu[ok]).all()
    assert np.allclose(gmu[ok], mu[ok], rtol=rtol, atol=1e-10), np.abs(gmu[ok] - mu[ok]).max()
    assert np.allclose(gvar[ok], var[ok], rtol=rtol, atol=1e-10), np.abs(gvar[ok] - var[ok]).max()


def test_reference_missing_data_case(gsm):
    """test/krig.jl:284-312 through fit / predictprob: SK, OK, DK ~ 0, 1 < UK < 2, variances >= 0."""
    P = gsm.PointSet(np.array([[25.0, 25.0], [50.0, 75.0], [75.0, 50.0]]))
    data = gsm.georef({"z": [None, 2.0, 3.0]}, P)
    g = gsm.GaussianVariogram()
    x0 = np.array([50.0, 50.0])
    for model in (gsm.SimpleKriging(g, 0.0), gsm.OrdinaryKriging(g), gsm.UniversalKriging(g, exps=[(0, 0)])):
        d = gsm.predictprob(gsm.fit(model, data), "z", x0)
        assert abs(d.mu) < 1e-6 and d.sigma ** 2 >= 0


def test_benchmark_kriging_miss_shape(gsm, oracle):
    """benchmark/benchmarks.jl:26: OK GaussianVariogram(range=35), z = [missing, 2, 3], 100 x 100 grid, neighbors=true."""
    o = oracle
    X = np.array([[25.0, 25.0], [50.0, 75.0], [75.0, 50.0]])
    z = np.array([np.nan, 2.0, 3.0])
    got = gsm.fitpredict(gsm.OrdinaryKriging(gsm.GaussianVariogram(range=35.0)), gsm.georef({"z": z}, gsm.PointSet(X)),
                         gsm.CartesianGrid(100, 100), prob=True).z
    mu, var, st = o.fitpredict(o.ordinary_kriging(o.Variogram(o.GAUSSIAN, 1.0, 0.0, 35.0)), X, z,
                               o.Grid((0.0, 0.0), (1.0, 1.0), (100, 100)), prob=True)
    _cmp(got, mu, var, st)


@pytest.mark.parametrize("kind,deg,dim,k,frac", [("ok", None, 2, 10, 0.1), ("sk", None, 2, 16, 0.3), ("uk", 1, 2, 12, 0.15),
                                                  ("uk", 2, 2, 24, 0.1), ("ok", None, 3, 32, 0.05), ("uk", 1, 3, 20, 0.5),
                                                  ("uk", 2, 3, 32, 0.08), ("ok", None, 2, 32, 0.9)])
def test_local_kriging_with_missing_values(gsm, oracle, kind, deg, dim, k, frac):
    o = oracle
    rng = np.random.default_rng(50 + k + dim)
    n = 1500
    dims = (28, 19) if dim == 2 else (9, 8, 7)
    X = rng.uniform(0, 1, size=(n, dim)) * np.array(dims, dtype=float)
    z = np.sin(X[:, 0] / 5) + 0.1 * rng.standard_normal(n)
    z[rng.random(n) < frac] = np.nan
    vario = o.Variogram(o.SPHERICAL, 1.0, 0.1, 6.0)
    fun = gsm.SphericalVariogram(sill=1.0, range=6.0, nugget=0.1)
    zmean = float(np.nanmean(z))
    om = {"sk": lambda: o.simple_kriging(vario, zmean), "ok": lambda: o.ordinary_kriging(vario),
          "uk": lambda: o.universal_kriging(vario, deg, dim)}[kind]()
    gm = {"sk": lambda: gsm.SimpleKriging(fun, zmean), "ok": lambda: gsm.OrdinaryKriging(fun),
          "uk": lambda: gsm.UniversalKriging(fun, deg, dim)}[kind]()
    og = o.Grid(tuple(0.0 for _ in dims), tuple(1.0 for _ in dims), dims)
    mu, var, st = o.fitpredict(om, X, z, og, prob=True, maxneighbors=k)
    try:
        got = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), gsm.CartesianGrid(*dims), maxneighbors=k, prob=True).z
    except ValueError:        # a non-positive variance somewhere: the oracle must see it too
        assert (var[st == 0] <= 0).any()
        return
    # targets whose neighbourhood keeps (lambda_n, nu) unique; all-missing neighbourhoods (no sample with a value) give
    # mean 0 / variance sill on both sides
    gmu = np.ma.filled(got.mu, np.nan) if np.ma.isMaskedArray(got.mu) else np.asarray(got.mu)
    sing = ~np.isfinite(gmu)
    assert sing.mean() < 0.05
    st2 = np.where(sing, 2, st)
    _cmp(got, mu, var, st2, rtol=1e-8 if kind == "uk" else 1e-9)


def test_global_kriging_with_missing_values(gsm, oracle):
    o = oracle
    rng = np.random.default_rng(7)
    X = rng.uniform(0, 40, size=(300, 2))
    z = np.sin(X[:, 0] / 6) + 0.1 * rng.standard_normal(300)
    z[rng.random(300) < 0.2] = np.nan
    vario = o.Variogram(o.EXPONENTIAL, 1.0, 0.05, 12.0)
    fun = gsm.ExponentialVariogram(sill=1.0, range=12.0, nugget=0.05)
    og = o.Grid((0.0, 0.0), (2.0, 2.0), (20, 20))
    grid = gsm.CartesianGrid((0.0, 0.0), (40.0, 40.0), dims=(20, 20))
    for om, gm in ((o.ordinary_kriging(vario), gsm.OrdinaryKriging(fun)), (o.simple_kriging(vario, 0.3), gsm.SimpleKriging(fun, 0.3)),
                   (o.universal_kriging(vario, 1, 2), gsm.UniversalKriging(fun, 1, 2))):
        mu, var, st = o.fitpredict(om, X, z, og, prob=True, neighbors=False)
        got = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), grid, neighbors=False, prob=True).z
        assert np.allclose(got.mu, mu, rtol=1e-8, atol=1e-10) and np.allclose(got.sigma, var, rtol=1e-8, atol=1e-10)


def test_dirac_and_nn_ball(gsm, oracle):
    rng = np.random.default_rng(3)
    X = rng.uniform(0, 30, size=(200, 2))
    z = rng.standard_normal(200)
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    grid = gsm.CartesianGrid(30, 30)
    a = gsm.fitpredict(gsm.IDW(2), tab, grid, maxneighbors=8).z
    b = gsm.fitpredict(gsm.IDW(2), tab, grid, maxneighbors=8, prob=True).z
    assert isinstance(b, gsm.Dirac) and np.array_equal(np.asarray(a), np.asarray(b.value))
    d = gsm.predictprob(gsm.fit(gsm.IDW(2), tab), "z", np.array([3.0, 4.0]))
    assert isinstance(d, gsm.Dirac) and d.value == gsm.predict(gsm.fit(gsm.IDW(2), tab), "z", np.array([3.0, 4.0]))
    # NN with a ball neighbourhood: fewer than minneighbors samples inside the ball -> missing (models.jl:167-181)
    c = gsm.fitpredict(gsm.NN(), tab, grid, neighborhood=1.0, minneighbors=2, maxneighbors=5).z
    T = np.stack(np.meshgrid(np.arange(30) + 0.5, np.arange(30) + 0.5, indexing="xy"), axis=-1).reshape(-1, 2)
    d2 = ((T[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    inside = (np.sqrt(d2) <= 1.0).sum(1)
    assert np.array_equal(np.ma.getmaskarray(c), np.minimum(inside, 5) < 2)


def test_idw_with_missing_values(gsm, oracle):
    """idw.jl:59-73 on a column with `missing`: the weighted sum over samples that include a missing value is `missing`;
    an exact hit returns that sample's value (missing or not) whatever the other samples hold."""
    o = oracle
    rng = np.random.default_rng(17)
    X = rng.uniform(0, 30, size=(2000, 2))
    z = np.sin(X[:, 0] / 5) + 0.1 * rng.standard_normal(2000)
    z[rng.random(2000) < 0.02] = np.nan
    P = np.concatenate([rng.uniform(0, 30, size=(400, 2)), X[:40]])          # 40 exact hits, some on missing samples
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    got = gsm.fitpredict(gsm.IDW(2), tab, gsm.PointSet(P), maxneighbors=8)
    ref, _, st = o.fitpredict(o.IDWModel(2), X, z, P, maxneighbors=8)
    col = got.z
    assert np.ma.isMaskedArray(col) and np.array_equal(np.ma.getmaskarray(col), np.isnan(ref))
    assert 0 < np.isnan(ref).sum() < len(ref)
    assert np.allclose(np.ma.getdata(col)[~np.isnan(ref)], ref[~np.isnan(ref)], rtol=1e-12, atol=1e-13)
    hits = np.ma.getdata(col)[400:]
    assert np.array_equal(np.isnan(z[:40]), np.ma.getmaskarray(col)[400:]) and np.array_equal(hits[~np.isnan(z[:40])], z[:40][~np.isnan(z[:40])])
    # all samples (neighbors=false): every prediction away from a sample with a value is missing
    got = gsm.fitpredict(gsm.IDW(1), tab, gsm.PointSet(P), neighbors=False).z
    ref, _, _ = o.fitpredict(o.IDWModel(1), X, z, P, neighbors=False)
    assert np.array_equal(np.ma.getmaskarray(got), np.isnan(ref)) and np.isnan(ref[:400]).all()
    assert np.array_equal(np.ma.getdata(got)[~np.isnan(ref)], ref[~np.isnan(ref)])
    f = gsm.fit(gsm.IDW(1), tab)
    assert gsm.predict(f, "z", P[0]) is None
    j = int(np.flatnonzero(~np.isnan(z[:40]))[0])
    assert gsm.predict(f, "z", X[j]) == z[j]
