"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200.

Tolerances (BASELINE.json north_star): neighbour index lists bit-exact under the total order
(d2, row); means and variances within relative 1e-10 in Float64.  `atol` below is the absolute floor
for quantities that are differences of O(sill) terms (an exact-hit variance is 1e-10 + rounding)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-10
ATOL = 1e-12


def _samples(seed, n, dim, extent):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, extent, size=(n, dim))
    z = np.sin(X[:, 0] / 20) + np.cos(X[:, -1] / 30) + 0.1 * rng.standard_normal(n)
    return X, z


def _grid_targets(gsm, o, dims, alpha=1.0, origin=None, spacing=None):
    origin = origin or tuple(0.0 for _ in dims)
    spacing = spacing or tuple(1.0 for _ in dims)
    og = o.Grid(origin, spacing, dims).scaled(alpha) if alpha != 1.0 else o.Grid(origin, spacing, dims)
    t = gsm.Context.grid_targets(og.origin, og.spacing, dims)
    return t, og


# ------------------------------------------------------------------------------------------ kNN
@pytest.mark.parametrize("dim,n,k", [(2, 5000, 32), (3, 8000, 32), (2, 300, 8), (3, 4000, 64), (1, 500, 16)])
def test_knn_lists_bit_exact(gsm, oracle, ctx, dim, n, k):
    X, z = _samples(10 + dim, n, dim, 64.0)
    ctx.set_samples(X, z)
    dims = {1: (200,), 2: (40, 30), 3: (12, 10, 9)}[dim]
    sp = tuple(64.0 / d for d in dims)
    t, og = _grid_targets(gsm, oracle, dims, spacing=sp)
    idx, cnt = ctx.knn(t, k)
    T = og.centroids()
    assert (cnt == k).all()
    for j in range(T.shape[0]):
        ref = oracle.knn(X, T[j], k)
        assert np.array_equal(idx[j], ref), (j, idx[j], ref)


def test_knn_ties_on_lattice_samples(gsm, oracle, ctx):
    """Samples on an integer lattice, targets on lattice points and cell centres: massive distance ties."""
    g = np.stack(np.meshgrid(np.arange(24.0), np.arange(24.0), indexing="ij"), -1).reshape(-1, 2)
    rng = np.random.default_rng(5)
    X = g[rng.permutation(len(g))]
    z = rng.standard_normal(len(X))
    ctx.set_samples(X, z)
    P = np.concatenate([g[::7], g[::11] + 0.5])
    t = gsm.Context.point_targets(P)
    idx, cnt = ctx.knn(t, 16)
    for j in range(len(P)):
        assert np.array_equal(idx[j], oracle.knn(X, P[j], 16)), j


def test_knn_targets_outside_bbox_and_ball(gsm, oracle, ctx):
    X, z = _samples(3, 2000, 2, 50.0)
    ctx.set_samples(X, z)
    P = np.array([[-30.0, -30.0], [80.0, 10.0], [25.0, 200.0], [25.0, 25.0], [0.0, 0.0]])
    idx, cnt = ctx.knn(gsm.Context.point_targets(P), 12)
    for j in range(len(P)):
        assert np.array_equal(idx[j], oracle.knn(X, P[j], 12))
    idx, cnt = ctx.knn(gsm.Context.point_targets(P), 12, radius=2.0)   # KBallSearch
    for j in range(len(P)):
        ref = oracle.knn(X, P[j], 12, radius=2.0)
        assert cnt[j] == len(ref)
        assert np.array_equal(idx[j, :cnt[j]], ref)
        assert (idx[j, cnt[j]:] == -1).all()


def test_knn_fewer_samples_than_k(gsm, oracle, ctx):
    X, z = _samples(4, 5, 2, 10.0)
    ctx.set_samples(X, z)
    P = np.array([[1.0, 1.0], [9.0, 2.0]])
    idx, cnt = ctx.knn(gsm.Context.point_targets(P), 32)   # clamps to N (models.jl:144-147)
    assert idx.shape == (2, 5) and (cnt == 5).all()
    for j in range(2):
        assert np.array_equal(idx[j], oracle.knn(X, P[j], 5))


# --------------------------------------------------------------------------- local Kriging parity
def _oracle_model(o, kind, vario, dim, mean=0.0, deg=None):
    if kind == "sk":
        return o.simple_kriging(vario, mean)
    if kind == "ok":
        return o.ordinary_kriging(vario)
    return o.universal_kriging(vario, deg, dim)


def _gsm_model(gsm, kind, vario, dim, mean=0.0, deg=None):
    cls = {0: gsm.GaussianVariogram, 1: gsm.SphericalVariogram, 2: gsm.ExponentialVariogram}[vario.kind]
    fun = cls(sill=vario.sill, range=vario.range, nugget=vario.nugget)
    if kind == "sk":
        return gsm.SimpleKriging(fun, mean)
    if kind == "ok":
        return gsm.OrdinaryKriging(fun)
    return gsm.UniversalKriging(fun, deg, dim)


CASES = [
    # kind, vario kind, dim, deg, k
    ("ok", 1, 2, None, 32),
    ("sk", 1, 2, None, 32),
    ("ok", 2, 2, None, 16),
    ("ok", 0, 2, None, 12),
    ("uk", 1, 2, 1, 32),
    ("uk", 1, 2, 2, 32),
    ("ok", 1, 3, None, 32),
    ("uk", 1, 3, 1, 32),
    ("uk", 1, 3, 2, 32),
    ("sk", 2, 3, None, 8),
    ("ok", 1, 2, None, 64),     # 32 < k <= 64: CTA-per-system capacity path
    ("uk", 1, 3, 1, 48),
    ("sk", 0, 2, None, 40),
]


@pytest.mark.parametrize("kind,vk,dim,deg,k", CASES)
def test_local_kriging_matches_oracle(gsm, oracle, kind, vk, dim, deg, k):
    o = oracle
    n = 3000 if dim == 2 else 6000
    X, z = _samples(100 + vk + dim, n, dim, 100.0)
    vario = o.Variogram(vk, 1.0, 0.1, 25.0)
    om = _oracle_model(o, kind, vario, dim, mean=float(z.mean()), deg=deg)
    gm = _gsm_model(gsm, kind, vario, dim, mean=float(z.mean()), deg=deg)
    dims = (24, 20) if dim == 2 else (8, 7, 6)
    sp = tuple(100.0 / d for d in dims)
    mu, var, st = o.fitpredict(om, X, z, o.Grid(tuple(0.0 for _ in dims), sp, dims), maxneighbors=k, prob=True)
    grid = gsm.CartesianGrid(tuple(0.0 for _ in dims), tuple(100.0 for _ in dims), dims=dims)
    pred = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), grid, maxneighbors=k, prob=True)
    got = pred.z
    assert (st == 0).all()
    assert np.allclose(got.mu, mu, rtol=RTOL, atol=ATOL), np.abs(got.mu - mu).max()
    assert np.allclose(got.sigma, var, rtol=RTOL, atol=ATOL), np.abs(got.sigma - var).max()
    pred2 = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), grid, maxneighbors=k, prob=False)
    assert np.array_equal(pred2.z, got.mu)


def test_local_kriging_golden_misc_jl(gsm):
    """reference test/misc.jl:13-25 through the GPU path (neighbors=true; N=3 < maxneighbors)."""
    data = gsm.georef({"z": np.array([1.0, 0.0, 1.0])}, [(25.0, 25.0), (50.0, 75.0), (75.0, 50.0)])
    grid = gsm.CartesianGrid((0.5, 0.5), (100.5, 100.5), dims=(100, 100))
    pred = gsm.fitpredict(gsm.Kriging(gsm.SphericalVariogram(range=35.0)), data, grid, neighbors=True)
    lin = lambda i, j: (i - 1) + (j - 1) * 100
    assert pred.z[lin(25, 25)] == pytest.approx(1.0, abs=1e-3)
    assert pred.z[lin(50, 75)] == pytest.approx(0.0, abs=1e-3)
    assert pred.z[lin(75, 50)] == pytest.approx(1.0, abs=1e-3)
    assert pred.geometry == grid   # test/misc.jl:53-59


def test_local_kriging_golden_docstrings(gsm, ctx):
    """reference test/krig.jl:451-479 known answers, via the local kernel with all 3 samples as neighbours
    (no rescaling: fit/predict are called directly in the reference test)."""
    L = gsm._lib
    X = np.array([[25.0, 25.0], [50.0, 75.0], [75.0, 50.0]])
    t = gsm.Context.point_targets(np.array([[50.0, 50.0]]))
    sph = L.make_variogram(L.GSM_VARIO_SPHERICAL, 1.0, 0.0, 1.0)
    ctx.set_samples(X, np.array([1.0, 0.0, 0.0]))
    mean, _, _, _ = ctx.local_krige(sph, L.make_model(L.GSM_KRIG_SIMPLE, mean=5.0), t, 3)
    assert mean[0] == pytest.approx(5.0, rel=1e-14)
    mean, _, _, _ = ctx.local_krige(sph, L.make_model(L.GSM_KRIG_ORDINARY), t, 3)
    assert mean[0] == pytest.approx(1 / 3, rel=1e-14)
    mean, _, _, _ = ctx.local_krige(sph, L.make_model(L.GSM_KRIG_UNIVERSAL, exps=[(0, 0), (1, 0), (0, 1)]), t, 3)
    assert mean[0] == pytest.approx(1 / 3, rel=1e-12)
    uk2 = L.make_model(L.GSM_KRIG_UNIVERSAL, exps=[(0, 0), (2, 0)])   # not a lower set: raw monomials
    mean, _, _, _ = ctx.local_krige(sph, uk2, t, 3)
    assert mean[0] == pytest.approx(0.4081632653061225, rel=1e-12)
    ctx.set_samples(X, np.array([0.0, 1.0, 0.0]))
    mean, _, _, _ = ctx.local_krige(sph, uk2, t, 3)
    assert mean[0] == pytest.approx(0.34693877551020413, rel=1e-12)


def test_local_kriging_properties(gsm, oracle):
    """Property tests of reference test/krig.jl:29-126 restated against the GPU library."""
    X, z = _samples(77, 4000, 2, 100.0)
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    fun = gsm.SphericalVariogram(range=30.0, sill=1.0, nugget=0.0)
    sk = gsm.SimpleKriging(fun, float(z.mean()))
    ok = gsm.OrdinaryKriging(fun)
    # interpolation at sample locations, var >= 0, var(SK) <= var(OK)
    P = gsm.PointSet(X[:200])
    psk = gsm.fitpredict(sk, tab, P, maxneighbors=32, prob=True).z
    pok = gsm.fitpredict(ok, tab, P, maxneighbors=32, prob=True).z
    puk = gsm.fitpredict(gsm.UniversalKriging(fun, 1, 2), tab, P, maxneighbors=32, prob=True).z
    for p in (psk, pok, puk):
        assert np.allclose(p.mu, z[:200], rtol=1e-9, atol=1e-9)
        assert (p.sigma >= 0).all()
    assert (psk.sigma <= pok.sigma + 1e-12).all()
    # OK == UK(deg 0)
    Q = gsm.CartesianGrid((0.0, 0.0), (100.0, 100.0), dims=(16, 16))
    a = gsm.fitpredict(ok, tab, Q, maxneighbors=32, prob=True).z
    b = gsm.fitpredict(gsm.UniversalKriging(fun, 0, 2), tab, Q, maxneighbors=32, prob=True).z
    assert np.allclose(a.mu, b.mu, rtol=1e-12) and np.allclose(a.sigma, b.sigma, rtol=1e-12)
    # sill scaling: mean unchanged, variance (minus the 1e-10 floor) scales
    fun2 = gsm.SphericalVariogram(range=30.0, sill=2.0, nugget=0.0)
    c = gsm.fitpredict(gsm.OrdinaryKriging(fun2), tab, Q, maxneighbors=32, prob=True).z
    assert np.allclose(c.mu, a.mu, rtol=1e-10)
    assert np.allclose(c.sigma - 1e-10, 2.0 * (a.sigma - 1e-10), rtol=1e-9, atol=1e-12)
    # variance is a function of the configuration, not of the values
    tab2 = gsm.georef({"z": z + np.random.default_rng(1).uniform(size=len(z))}, gsm.PointSet(X))
    d = gsm.fitpredict(ok, tab2, Q, maxneighbors=32, prob=True).z
    assert np.array_equal(d.sigma, a.sigma)


def test_minneighbors_missing_and_shards(gsm, oracle):
    X, z = _samples(8, 1500, 2, 60.0)
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    ok = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=20.0, nugget=0.05))
    P = np.concatenate([np.random.default_rng(2).uniform(0, 60, size=(64, 2)), [[500.0, 500.0]]])
    pred = gsm.fitpredict(ok, tab, gsm.PointSet(P), maxneighbors=16, minneighbors=3, neighborhood=6.0, prob=True)
    om = oracle.ordinary_kriging(oracle.Variogram(oracle.SPHERICAL, 1.0, 0.05, 20.0))
    mu, var, st = oracle.fitpredict(om, X, z, P, maxneighbors=16, minneighbors=3, neighborhood=6.0, prob=True)
    got = pred.z
    miss = np.ma.getmaskarray(got.mu)
    assert np.array_equal(miss, st == 1) and miss[-1]
    assert np.allclose(np.ma.filled(got.mu, 0)[~miss], mu[~miss], rtol=RTOL, atol=ATOL)
    assert np.allclose(np.ma.filled(got.sigma, 0)[~miss], var[~miss], rtol=RTOL, atol=ATOL)
    # the device-side summary (gsm_get_summary) agrees with the status column
    c = gsm.default_context()
    sm = c.summary()
    assert sm["total"] == st.size and sm["missing"] == int((st == 1).sum()) and sm["singular"] == int((st == 2).sum())
    assert sm["nonpositive_var"] == 0
    # shards of a grid concatenate to the whole domain
    grid = gsm.CartesianGrid((0.0, 0.0), (60.0, 60.0), dims=(20, 15))
    whole = gsm.fitpredict(ok, tab, grid, maxneighbors=16, prob=True).z
    parts = [gsm.fitpredict(ok, tab, grid, maxneighbors=16, prob=True, first=f, count=c).z
             for f, c in ((0, 100), (100, 77), (177, 123))]
    # every system is assembled in a canonical order (samples shared by its 4 x 2 patch first, then ascending
    # record position); a shard boundary that cuts a patch can change which samples count as shared, so the
    # guarantee is agreement to rounding, not bit-equality
    assert np.allclose(np.concatenate([p.mu for p in parts]), whole.mu, rtol=1e-12, atol=1e-13)
    assert np.allclose(np.concatenate([p.sigma for p in parts]), whole.sigma, rtol=1e-12, atol=1e-13)


# ------------------------------------------------------------------------------------------- IDW
@pytest.mark.parametrize("exponent", [1, 2, 3, 2.5])
def test_idw_global_matches_oracle(gsm, oracle, exponent):
    X, z = _samples(21, 2000, 2, 256.0)
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    grid = gsm.CartesianGrid(24, 24)
    pred = gsm.fitpredict(gsm.IDW(exponent), tab, grid, neighbors=False).z
    mu, _, _ = oracle.fitpredict(oracle.IDWModel(exponent), X, z, oracle.Grid((0.0, 0.0), (1.0, 1.0), (24, 24)), neighbors=False)
    assert np.allclose(pred, mu, rtol=RTOL, atol=ATOL), np.abs(pred - mu).max()


def test_idw_local_and_exact_hits(gsm, oracle):
    X, z = _samples(22, 3000, 3, 40.0)
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    P = np.concatenate([X[:50], np.random.default_rng(9).uniform(0, 40, size=(200, 3))])
    pred = gsm.fitpredict(gsm.IDW(2), tab, gsm.PointSet(P), maxneighbors=10).z
    mu, _, _ = oracle.fitpredict(oracle.IDWModel(2), X, z, P, maxneighbors=10)
    assert np.allclose(pred, mu, rtol=RTOL, atol=ATOL)
    assert np.array_equal(pred[:50], z[:50])   # exact hits return the sample value (test/idw.jl:9-16)
    # test/misc.jl:2-8: IDW on its own point set reproduces the data exactly
    data = gsm.georef({"z": np.array([1.0, 2.0, 3.0])}, [(25.0, 25.0), (50.0, 75.0), (75.0, 25.0)])
    out = gsm.fitpredict(gsm.IDW(), data, data.domain, neighbors=False)
    assert np.array_equal(out.z, data.z) and out.geometry == data.geometry


def test_error_codes(gsm):
    L = gsm._lib
    c = gsm.Context(0)
    t = gsm.Context.point_targets(np.zeros((1, 2)))
    with pytest.raises(gsm.GsmError) as e:
        c.idw(2.0, t)
    assert e.value.code == L.GSM_ERR_NO_SAMPLES
    c.set_samples(np.random.default_rng(0).uniform(size=(100, 2)), np.zeros(100))
    with pytest.raises(gsm.GsmError) as e:
        c.local_krige(L.make_variogram(9, 1, 0, 1), L.make_model(L.GSM_KRIG_ORDINARY), t, 8)
    assert e.value.code == L.GSM_ERR_INVALID
    with pytest.raises(gsm.GsmError) as e:
        c.local_krige(L.make_variogram(1, 1, 0, 1), L.make_model(L.GSM_KRIG_ORDINARY),
                      gsm.Context.point_targets(np.zeros((1, 3))), 8)
    assert e.value.code == L.GSM_ERR_INVALID
    c.close()


# ------------------------------------------------------------------------------- global Kriging
GLOBAL_CASES = [("ok", 1, 2, None, 700), ("sk", 1, 2, None, 300), ("uk", 1, 2, 1, 500), ("uk", 2, 3, 2, 400),
                ("ok", 2, 3, None, 129), ("ok", 1, 1, None, 64)]


@pytest.mark.parametrize("kind,vk,dim,deg,n", GLOBAL_CASES)
def test_global_kriging_matches_oracle(gsm, oracle, kind, vk, dim, deg, n):
    o = oracle
    X, z = _samples(200 + vk + dim, n, dim, 100.0)
    vario = o.Variogram(vk, 1.0, 0.1, 30.0)
    om = _oracle_model(o, kind, vario, dim, mean=float(z.mean()), deg=deg)
    gm = _gsm_model(gsm, kind, vario, dim, mean=float(z.mean()), deg=deg)
    dims = {1: (150,), 2: (13, 11), 3: (6, 5, 5)}[dim]
    sp = tuple(100.0 / d for d in dims)
    mu, var, st = o.fitpredict(om, X, z, o.Grid(tuple(0.0 for _ in dims), sp, dims), neighbors=False, prob=True)
    grid = gsm.CartesianGrid(tuple(0.0 for _ in dims), tuple(100.0 for _ in dims), dims=dims)
    got = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), grid, neighbors=False, prob=True).z
    assert np.allclose(got.mu, mu, rtol=RTOL, atol=ATOL), np.abs(got.mu - mu).max()
    assert np.allclose(got.sigma, var, rtol=RTOL, atol=ATOL), np.abs(got.sigma - var).max()


def test_global_kriging_golden_misc_jl(gsm):
    """reference test/misc.jl:13-21 (neighbors=false) through fit-once / predict-all on the GPU."""
    data = gsm.georef({"z": np.array([1.0, 0.0, 1.0])}, [(25.0, 25.0), (50.0, 75.0), (75.0, 50.0)])
    grid = gsm.CartesianGrid((0.5, 0.5), (100.5, 100.5), dims=(100, 100))
    pred = gsm.fitpredict(gsm.Kriging(gsm.SphericalVariogram(range=35.0)), data, grid, neighbors=False)
    lin = lambda i, j: (i - 1) + (j - 1) * 100
    assert pred.z[lin(25, 25)] == pytest.approx(1.0, abs=1e-3)
    assert pred.z[lin(50, 75)] == pytest.approx(0.0, abs=1e-3)
    assert pred.z[lin(75, 50)] == pytest.approx(1.0, abs=1e-3)


def test_fit_predict_predictprob_status_api(gsm):
    """The single-target API of reference test/krig.jl:431-479 (fit / predict / predictprob / status)."""
    pset = gsm.PointSet([(25.0, 25.0), (50.0, 75.0), (75.0, 50.0)])
    data = gsm.georef({"a": np.array([1.0, 0.0, 0.0])}, pset)
    p = (50.0, 50.0)
    fm = gsm.fit(gsm.Kriging(gsm.SphericalVariogram(), 5.0), data)
    assert gsm.predict(fm, "a", p) == pytest.approx(5.0, rel=1e-14)
    assert gsm.status(fm)
    fm = gsm.fit(gsm.Kriging(gsm.SphericalVariogram()), data)
    assert gsm.predict(fm, "a", p) == pytest.approx(1 / 3, rel=1e-14)
    assert gsm.predict(fm, ("a",), p) == [gsm.predict(fm, "a", p)]
    d = gsm.predictprob(fm, "a", p)
    assert isinstance(d, gsm.Normal) and d.mu == pytest.approx(1 / 3) and d.sigma == pytest.approx(np.sqrt(4 / 3 + 1e-10))
    with pytest.raises(ValueError):
        gsm.predict(fm, ("a", "b"), p)   # ArgumentError, models.jl:56-57
    fm = gsm.fit(gsm.Kriging(gsm.SphericalVariogram(), [(0, 0), (1, 0), (0, 1)]), data)
    assert gsm.predict(fm, "a", p) == pytest.approx(1 / 3, rel=1e-12)
    fm = gsm.fit(gsm.Kriging(gsm.SphericalVariogram(), [(0, 0), (2, 0)]), data)
    assert gsm.predict(fm, "a", p) == pytest.approx(0.4081632653061225, rel=1e-12)
    with pytest.raises(ValueError):   # checkcompat, krig.jl:78-85
        gsm.fit(gsm.Kriging(gsm.SphericalVariogram()), gsm.georef({"a": np.zeros(3), "b": np.zeros(3)}, pset))
    # interpolation property at the samples (krig.jl:29-47) for a Gaussian variogram
    rng = np.random.default_rng(2024)
    P = rng.uniform(0, 10, size=(10, 3))
    zz = rng.uniform(size=10)
    tab = gsm.georef({"z": zz}, gsm.PointSet(P))
    for model in (gsm.SimpleKriging(gsm.GaussianVariogram(), float(zz.mean())), gsm.OrdinaryKriging(gsm.GaussianVariogram()),
                  gsm.UniversalKriging(gsm.GaussianVariogram(), 1, 3)):
        fm = gsm.fit(model, tab)
        for j in range(10):
            d = gsm.predictprob(fm, "z", P[j])
            assert d.mu == pytest.approx(zz[j], abs=1e-5) and d.sigma >= 0
    # IDW fitted model (test/idw.jl:26-43)
    fi = gsm.fit(gsm.IDW(), gsm.georef({"z": np.array([1.0, 0.0, 1.0])}, [(0.0, 0.0), (1.0, 0.0), (1.0, 1.0)]))
    assert gsm.predict(fi, "z", (0.0, 0.0)) == 1.0 and gsm.status(fi)


def test_global_kriging_singular_status(gsm):
    """Duplicate samples with zero nugget: the factorisation fails and status(fitted) is false
    (krig.jl:49-52, check=false at krig.jl:152)."""
    P = np.array([[0.0, 0.0], [1.0, 1.0], [1.0, 1.0], [2.0, 0.5]])
    fm = gsm.fit(gsm.OrdinaryKriging(gsm.SphericalVariogram(range=5.0)), gsm.georef({"z": np.arange(4.0)}, gsm.PointSet(P)))
    assert not gsm.status(fm)


def test_uk_quadratic_3d_against_exact_arithmetic(gsm, oracle):
    """C4-class systems (3-D, k = 32, quadratic drift, dense samples).  The reference factors the saddle
    matrix with RAW monomials, whose conditioning costs it several digits; the GPU evaluates the same
    polynomial space centred at the target.  Against 50-digit arithmetic (mpmath) the GPU must hold
    1e-10 and be at rounding level or no worse than the float64 oracle."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    o = oracle
    rng = np.random.default_rng(4)
    n = 30000
    X = rng.uniform(0, 64, size=(n, 3))
    z = np.sin(X[:, 0] / 20) + np.cos(X[:, 2] / 30) + 0.1 * rng.standard_normal(n)
    P = rng.uniform(8, 56, size=(5, 3))
    vario = o.Variogram(o.SPHERICAL, 1.0, 0.1, 20.0)
    om = o.universal_kriging(vario, 2, 3)
    gm = gsm.UniversalKriging(gsm.SphericalVariogram(sill=1.0, range=20.0, nugget=0.1), 2, 3)
    got = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), gsm.PointSet(P), maxneighbors=32, prob=True).z
    mu_o, var_o, _ = o.fitpredict(om, X, z, P, maxneighbors=32, prob=True)
    alpha = o.scale_factor(vario.range, o.points_absmax(X), o.points_absmax(P))
    sv = vario.scaled(alpha)
    exps = o.monomial_exponents(2, 3)

    def cov(h):
        if h == 0:
            return mp.mpf(sv.sill)
        t = h / mp.mpf(sv.range)
        g = (mp.mpf(sv.sill) - mp.mpf(sv.nugget)) * (mp.mpf("1.5") * t - mp.mpf("0.5") * t ** 3) if h < sv.range \
            else mp.mpf(sv.sill) - mp.mpf(sv.nugget)
        return mp.mpf(sv.sill) - (g + mp.mpf(sv.nugget))

    for j in range(len(P)):
        x0 = P[j] * alpha
        idx = o.knn(X * alpha, x0, 32)
        Q = [[mp.mpf(float(c)) for c in row] for row in (X * alpha)[idx]]
        t0 = [mp.mpf(float(c)) for c in x0]
        k, p = 32, len(exps)
        A = mp.matrix(k + p, k + p)
        r = mp.matrix(k + p, 1)
        for a in range(k):
            for b in range(k):
                A[a, b] = cov(mp.sqrt(sum((Q[a][d] - Q[b][d]) ** 2 for d in range(3))))
            r[a] = cov(mp.sqrt(sum((Q[a][d] - t0[d]) ** 2 for d in range(3))))
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
            r[k + q] = f
        w = mp.lu_solve(A, r)
        mu_e = sum(w[a] * mp.mpf(float(z[idx[a]])) for a in range(k))
        var_e = cov(0) - sum(r[a] * w[a] for a in range(k + p)) + mp.mpf("1e-10")
        err_gpu = max(abs((mp.mpf(float(got.mu[j])) - mu_e) / mu_e), abs((mp.mpf(float(got.sigma[j])) - var_e) / var_e))
        err_ora = max(abs((mp.mpf(float(mu_o[j])) - mu_e) / mu_e), abs((mp.mpf(float(var_o[j])) - var_e) / var_e))
        assert err_gpu < 1e-10, (j, float(err_gpu))
        # at rounding level, or else no worse than the float64 oracle
        assert err_gpu <= max(err_ora, mp.mpf("1e-13")), (j, float(err_gpu), float(err_ora))


# ------------------------------------------------------------------------------------ edge cases
def test_edge_cases_small_and_degenerate_inputs(gsm, oracle):
    o = oracle
    ok = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=5.0, nugget=0.01))
    om = o.ordinary_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.01, 5.0))
    # a single sample: every prediction is that value (OK weights sum to one)
    one = gsm.georef({"z": np.array([3.5])}, gsm.PointSet([(1.0, 2.0)]))
    pred = gsm.fitpredict(ok, one, gsm.CartesianGrid(4, 3), maxneighbors=32, prob=True).z
    assert np.allclose(pred.mu, 3.5, rtol=1e-14) and (pred.sigma > 0).all()
    # two samples, maxneighbors=1 (nearest-sample Kriging), 1-D domain
    X = np.array([[0.0], [10.0]])
    z = np.array([1.0, 2.0])
    P = np.array([[1.0], [4.9], [5.1], [9.0]])
    got = gsm.fitpredict(ok, gsm.georef({"z": z}, gsm.PointSet(X)), gsm.PointSet(P), maxneighbors=1).z
    assert np.array_equal(got, np.array([1.0, 1.0, 2.0, 2.0]))
    # empty target shard
    t = gsm.Context.grid_targets((0.0, 0.0), (1.0, 1.0), (8, 8), first=10, count=0)
    c = gsm.Context(0)
    c.set_samples(np.random.default_rng(0).uniform(size=(50, 2)), np.zeros(50))
    L = gsm._lib
    mean, var, st, _ = c.local_krige(L.make_variogram(1, 1.0, 0.0, 1.0), L.make_model(L.GSM_KRIG_ORDINARY), t, 8)
    assert mean.shape == (0,)
    c.close()
    # collinear samples in 2-D (degenerate bounding box in y) and targets far outside the box
    Xl = np.stack([np.linspace(0, 20, 60), np.zeros(60)], 1)
    zl = np.sin(Xl[:, 0])
    Pl = np.array([[5.0, 0.0], [5.0, 3.0], [-40.0, 7.0], [100.0, -100.0]])
    got = gsm.fitpredict(ok, gsm.georef({"z": zl}, gsm.PointSet(Xl)), gsm.PointSet(Pl), maxneighbors=12, prob=True).z
    mu, var, _ = o.fitpredict(om, Xl, zl, Pl, maxneighbors=12, prob=True)
    assert np.allclose(got.mu, mu, rtol=RTOL, atol=ATOL) and np.allclose(got.sigma, var, rtol=RTOL, atol=ATOL)
    # duplicate samples with zero nugget: singular local systems surface as status, not as an exception
    Xd = np.array([[0.0, 0.0], [1.0, 1.0], [1.0, 1.0], [2.0, 0.5], [3.0, 3.0]])
    c = gsm.Context(0)
    c.set_samples(Xd, np.arange(5.0))
    mean, var, st, _ = c.local_krige(L.make_variogram(1, 1.0, 0.0, 10.0), L.make_model(L.GSM_KRIG_ORDINARY),
                                     gsm.Context.point_targets(np.array([[1.5, 1.0]])), 5)
    assert st[0] == L.GSM_ST_SINGULAR and np.isnan(mean[0])
    c.close()


def test_large_sample_parity_against_c_oracle(gsm, oracle):
    """200 k samples, 20 000 scattered targets (the size class of BASELINE config C3) against the C twin of
    the oracle: neighbour lists bit-exact, mean/variance within 1e-10."""
    import gsm_oracle_c as oc
    o = oracle
    rng = np.random.default_rng(3)
    n = 200_000
    X = rng.uniform(0, 2048, size=(n, 2))
    z = np.sin(X[:, 0] / 20) + np.cos(X[:, 1] / 30) + 0.1 * rng.standard_normal(n)
    P = rng.uniform(-5, 2053, size=(20_000, 2))
    om = o.ordinary_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.1, 50.0))
    gm = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=50.0, sill=1.0, nugget=0.1))
    mu, var, st, nbr, cnt = oc.fitpredict(om, X, z, P, prob=True, maxneighbors=32, nthreads=8, want_nbr=True)
    got = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), gsm.PointSet(P), maxneighbors=32, prob=True).z
    assert np.allclose(got.mu, mu, rtol=RTOL, atol=ATOL) and np.allclose(got.sigma, var, rtol=RTOL, atol=ATOL)
    alpha = o.scale_factor(50.0, o.points_absmax(X), o.points_absmax(P))
    c = gsm.Context(0)
    c.set_samples(X * alpha, z)
    idx, k = c.knn(gsm.Context.point_targets(P * alpha), 32)
    assert np.array_equal(idx, nbr)
    c.close()


@pytest.mark.parametrize("sill", [1e13, 1e-13, 3.7e5])
@pytest.mark.parametrize("k,dims", [(10, (64, 48)), (20, (40, 30)), (32, (96, 64)), (5, (30, 20)), (40, (30, 20))])
def test_sill_spans_orders_of_magnitude(gsm, oracle, sill, k, dims):
    """The pivot test is relative to the sill, and padded rows (k not a multiple of 8, e.g. the default 10) carry
    sill * identity: a variogram with sill 1e13 or 1e-13 must predict like the reference."""
    import gsm_oracle_c as oc
    o = oracle
    rng = np.random.default_rng(31 + k)
    X = rng.uniform(0, 1, size=(3000, 2)) * np.array(dims, dtype=float)
    z = 1e6 * (np.sin(X[:, 0] / 9) + 0.1 * rng.standard_normal(3000))
    vario = o.Variogram(o.SPHERICAL, sill, 0.1 * sill, 12.0)
    for kind in ("ok", "sk", "uk"):
        om = _oracle_model(o, kind, vario, 2, mean=float(z.mean()), deg=1)
        gm = _gsm_model(gsm, kind, vario, 2, mean=float(z.mean()), deg=1)
        og = o.Grid((0.0, 0.0), (1.0, 1.0), dims)
        if kind == "uk":   # referee: the exact (binary128) solution (the float64 oracle loses digits on raw monomials)
            _, _, st, mu, var, okq = oc.fitpredict_exact(om, X, z, og, maxneighbors=k, nthreads=4)
            assert okq.all()
        else:
            mu, var, st, _, _ = oc.fitpredict(om, X, z, og, prob=True, maxneighbors=k, nthreads=4)
        got = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), gsm.CartesianGrid(*dims), maxneighbors=k, prob=True).z
        assert (st == 0).all() and np.isfinite(np.asarray(got.mu)).all()
        assert np.allclose(got.mu, mu, rtol=RTOL, atol=ATOL * 1e6), (kind, np.abs(got.mu - mu).max())
        assert np.allclose(got.sigma, var, rtol=RTOL, atol=ATOL * max(sill, 1e-10)), (kind, np.abs(got.sigma - var).max())


# ------------------------------------------------------------ shared-factor kernel, steady state
STEADY = [
    # kind, dim, deg, k, dims, n, first, count, neighborhood
    ("ok", 2, None, 32, (512, 96), 9000, 0, None, None),       # 6144 patches: ~20 jobs per CTA
    ("ok", 2, None, 32, (203, 77), 5000, 317, 14001, None),    # odd extents, slab cutting patches
    ("sk", 2, None, 16, (301, 60), 4000, 0, None, None),       # two block columns
    ("uk", 2, 1, 24, (256, 64), 6000, 0, None, None),          # ragged lists (k not a multiple of 8), drift rows
    ("ok", 3, None, 32, (40, 36, 34), 30000, 0, None, None),   # 2 x 2 x 2 patches
    ("uk", 3, 2, 32, (36, 30, 28), 25000, 1000, 25000, None),  # two right-hand-side block rows, slab
    ("ok", 2, None, 32, (256, 64), 2500, 0, None, 6.0),        # radius-limited: ragged and missing systems
    ("ok", 2, None, 32, (64, 48), 60000, 0, None, None),       # samples much denser than targets: pool overflow
    ("sk", 2, None, 8, (256, 96), 5000, 0, None, None),        # one block column: the per-system pipelined kernel
    ("ok", 2, None, 16, (256, 96), 30000, 0, None, None),      # dense samples, two block columns: pipelined kernel
    ("uk", 3, 1, 8, (40, 36, 34), 20000, 0, None, None),       # 3-D, one block column, drift rows
    ("ok", 2, None, 40, (128, 96), 6000, 0, None, None),       # 32 < k <= 64: several systems per persistent warp
    ("uk", 3, 1, 64, (24, 24, 20), 30000, 0, None, None),
    ("sk", 2, None, 64, (160, 64), 3000, 500, 9000, 7.5),      # k = 64, slab, radius-limited (ragged, missing)
]


@pytest.mark.parametrize("impl", ["auto", "patch", "pipe"])
@pytest.mark.parametrize("kind,dim,deg,k,dims,n,first,count,nbh", STEADY)
def test_shared_factor_kernel_steady_state(gsm, oracle, kind, dim, deg, k, dims, n, first, count, nbh, impl):
    """Grids large enough that every CTA of the persistent kernels runs many jobs (the 3-stage shared-factor
    pipeline and the descriptor / pool streaming in steady state), against the C oracle -- through the default
    dispatch and with every local-Kriging kernel forced (gsm_set_option "local_impl")."""
    import gsm_oracle_c as oc
    if impl != "auto" and k > 32:
        pytest.skip("k > 32 has one kernel")
    o = oracle
    ext = tuple(float(d) for d in dims)
    rng = np.random.default_rng(77 + dim + k)
    X = rng.uniform(0, 1, size=(n, dim)) * np.array(ext)
    z = np.sin(X[:, 0] / 20) + np.cos(X[:, -1] / 30) + 0.1 * rng.standard_normal(n)
    vario = o.Variogram(o.SPHERICAL, 1.0, 0.1, 25.0)
    om = _oracle_model(o, kind, vario, dim, mean=float(z.mean()), deg=deg)
    gm = _gsm_model(gsm, kind, vario, dim, mean=float(z.mean()), deg=deg)
    og = o.Grid(tuple(0.0 for _ in dims), tuple(1.0 for _ in dims), dims)
    kw = dict(prob=True, maxneighbors=k, first=first, count=count)
    if deg is None:
        mu, var, st, _, _ = oc.fitpredict(om, X, z, og, nthreads=8, neighborhood=nbh, **kw)
    else:
        # UniversalKriging: the referee is the exact (binary128) solution of the reference's equations on the same
        # neighbour lists -- float64 Bunch-Kaufman on raw monomials is itself up to ~1e-8 away from it
        kq = {a: b for a, b in kw.items() if a != "prob"}
        _, _, st, mu, var, okq = oc.fitpredict_exact(om, X, z, og, nthreads=8, neighborhood=nbh, **kq)
        assert okq[st == 0].all()
    grid = gsm.CartesianGrid(*dims)
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    kw["count"] = -1 if count is None else count
    c = gsm.Context(0)
    c.set_option("local_impl", impl)
    try:
        got = gsm.fitpredict(gm, tab, grid, neighborhood=nbh, ctx=c, **kw).z
    finally:
        c.close()
    miss = np.ma.getmaskarray(got.mu) if np.ma.isMaskedArray(got.mu) else ~np.isfinite(got.mu)
    assert np.array_equal(miss, st != 0)
    ok = ~miss
    gmu, gvar = np.ma.filled(got.mu, 0.0), np.ma.filled(got.sigma, 0.0)
    assert np.allclose(gmu[ok], mu[ok], rtol=RTOL, atol=ATOL), np.abs(gmu[ok] - mu[ok]).max()
    assert np.allclose(gvar[ok], var[ok], rtol=RTOL, atol=ATOL), np.abs(gvar[ok] - var[ok]).max()


@pytest.mark.parametrize("dims,k", [((3, 2), 12), ((5, 1), 16), ((1, 9), 10), ((2, 2, 2), 20), ((7, 3, 1), 32),
                                    ((9, 5, 3), 32), ((4, 2), 32), ((33, 17), 9)])
def test_shared_factor_kernel_tiny_and_degenerate_grids(gsm, oracle, dims, k):
    """Grids smaller than / not aligned with the 4 x 2 and 2 x 2 x 2 patches, fewer jobs than CTAs."""
    o = oracle
    dim = len(dims)
    rng = np.random.default_rng(5 + sum(dims) + k)
    X = rng.uniform(0, 10, size=(400, dim))
    z = rng.standard_normal(400)
    vario = o.Variogram(o.SPHERICAL, 1.0, 0.1, 4.0)
    og = o.Grid(tuple(0.0 for _ in dims), tuple(10.0 / d for d in dims), dims)
    mu, var, st = o.fitpredict(o.ordinary_kriging(vario), X, z, og, maxneighbors=k, prob=True)
    gm = gsm.OrdinaryKriging(gsm.SphericalVariogram(sill=1.0, range=4.0, nugget=0.1))
    grid = gsm.CartesianGrid(tuple(0.0 for _ in dims), tuple(10.0 for _ in dims), dims=dims)
    got = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), grid, maxneighbors=k, prob=True).z
    assert (st == 0).all()
    assert np.allclose(got.mu, mu, rtol=RTOL, atol=ATOL) and np.allclose(got.sigma, var, rtol=RTOL, atol=ATOL)


def test_full_size_c3_properties_and_parity(gsm, oracle):
    """BASELINE config C3 at full size (200 k samples -> 2048 x 2048 grid, local OK, k = 32): size-independent
    properties over all 4.19 M targets (the OK weights sum to one, so an affine map of the data maps the means
    affinely and leaves the variances alone), and parity against the C oracle on a strided subsample."""
    import gsm_oracle_c as oc
    o = oracle
    rng = np.random.default_rng(3)
    n = 200_000
    X = rng.uniform(0, 2048, size=(n, 2))
    z = np.sin(X[:, 0] / 20) + np.cos(X[:, 1] / 30) + 0.1 * rng.standard_normal(n)
    gm = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=50.0, sill=1.0, nugget=0.1))
    grid = gsm.CartesianGrid(2048, 2048)
    p1 = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), grid, maxneighbors=32, prob=True).z
    mu1, var1 = np.array(p1.mu), np.array(p1.sigma)
    p2 = gsm.fitpredict(gm, gsm.georef({"z": 2.0 * z + 1.0}, gsm.PointSet(X)), grid, maxneighbors=32, prob=True).z
    assert np.isfinite(mu1).all() and (var1 > 0).all()
    assert np.array_equal(var1, p2.sigma)                                     # the variance never sees the data
    assert np.allclose(p2.mu, 2.0 * mu1 + 1.0, rtol=1e-12, atol=1e-12)
    assert var1.max() <= 1.0 + 0.1 + 1e-9                                     # bounded by sill (+ Lagrange term < nugget)
    om = o.ordinary_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.1, 50.0))
    idx = np.unique(np.linspace(0, 2048 * 2048 - 1, 3000).astype(np.int64))
    og = o.Grid((0.0, 0.0), (1.0, 1.0), (2048, 2048))
    T = og.centroids()[idx]

    class _Dom(np.ndarray):
        pass

    dom = np.asarray(T).view(_Dom)
    orig = o._dom_absmax
    o._dom_absmax = lambda d: og.bbox_absmax() if isinstance(d, _Dom) else orig(d)
    try:
        mu, var, st, _, _ = oc.fitpredict(om, X, z, dom, prob=True, maxneighbors=32, nthreads=8)
    finally:
        o._dom_absmax = orig
    assert (st == 0).all()
    assert np.allclose(mu1[idx], mu, rtol=RTOL, atol=ATOL) and np.allclose(var1[idx], var, rtol=RTOL, atol=ATOL)


def test_full_size_north_star_3d(gsm, oracle):
    """The north-star target of BASELINE.json: local OK, k = 32, 1 M 3-D samples -> 256^3 grid, at full size:
    affine-data property over all 16.7 M targets and parity against the C oracle on a strided subsample."""
    import gsm_oracle_c as oc
    o = oracle
    rng = np.random.default_rng(6)
    n = 1_000_000
    X = rng.uniform(0, 256, size=(n, 3))
    z = np.sin(X[:, 0] / 20) + np.cos(X[:, 2] / 30) + 0.1 * rng.standard_normal(n)
    gm = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=20.0, sill=1.0, nugget=0.1))
    grid = gsm.CartesianGrid(256, 256, 256)
    p1 = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), grid, maxneighbors=32, prob=True).z
    mu1, var1 = np.array(p1.mu), np.array(p1.sigma)
    p2 = gsm.fitpredict(gm, gsm.georef({"z": 0.5 * z - 3.0}, gsm.PointSet(X)), grid, maxneighbors=32, prob=True).z
    assert np.isfinite(mu1).all() and (var1 > 0).all()
    assert np.array_equal(var1, p2.sigma)
    assert np.allclose(p2.mu, 0.5 * mu1 - 3.0, rtol=1e-12, atol=1e-12)
    om = o.ordinary_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.1, 20.0))
    idx = np.unique(np.linspace(0, 256 ** 3 - 1, 1500).astype(np.int64))
    og = o.Grid((0.0,) * 3, (1.0,) * 3, (256, 256, 256))
    T = np.stack([og.centroids(int(i), 1)[0] for i in idx])

    class _Dom(np.ndarray):
        pass

    dom = np.asarray(T).view(_Dom)
    orig = o._dom_absmax
    o._dom_absmax = lambda d: og.bbox_absmax() if isinstance(d, _Dom) else orig(d)
    try:
        mu, var, st, _, _ = oc.fitpredict(om, X, z, dom, prob=True, maxneighbors=32, nthreads=8)
    finally:
        o._dom_absmax = orig
    assert (st == 0).all()
    assert np.allclose(mu1[idx], mu, rtol=RTOL, atol=ATOL) and np.allclose(var1[idx], var, rtol=RTOL, atol=ATOL)


def _exact_report(name, got_mu, got_var, r):
    """|GPU - exact| and |float64 oracle - exact| in the allclose metric err / (atol + rtol |exact|) <= 1."""
    import json, os
    def score(a, e):
        return float(np.max(np.abs(a - e) / (ATOL + RTOL * np.abs(e))))
    def rel(a, e):
        return float(np.max(np.abs(a - e) / np.maximum(np.abs(e), 1e-300)))
    rep = {"config": name, "targets_checked": int(len(r["mu_q"])), "rtol": RTOL, "atol": ATOL,
           "gpu_vs_exact": {"mean_score": score(got_mu, r["mu_q"]), "var_score": score(got_var, r["var_q"]),
                            "mean_max_abs": float(np.abs(got_mu - r["mu_q"]).max()),
                            "var_max_rel": rel(got_var, r["var_q"])},
           "float64_oracle_vs_exact": {"mean_score": score(r["mu64"], r["mu_q"]), "var_score": score(r["var64"], r["var_q"]),
                                       "mean_max_abs": float(np.abs(r["mu64"] - r["mu_q"]).max()),
                                       "var_max_rel": rel(r["var64"], r["var_q"])},
           "gpu_vs_float64_oracle": {"mean_max_abs": float(np.abs(got_mu - r["mu64"]).max()),
                                     "var_max_rel": rel(got_var, r["var64"])},
           "pivot_ratio_max": float(r["pivot_ratio"].max())}
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "exact_" + name + ".json"), "w") as f:
        json.dump(rep, f, indent=1)
    return rep


def test_full_size_c4_uk_quadratic_against_exact_arithmetic(gsm, oracle):
    """BASELINE config C4 at full size: UniversalKriging, quadratic drift (10 monomials), k = 32, 1 M 3-D samples ->
    256^3 grid.  The reference factors the saddle matrix with RAW monomials of alpha-scaled coordinates, which costs
    float64 several digits, so GPU-vs-float64-oracle cannot meet 1e-10 by construction of the ORACLE.  The referee is
    the binary128 evaluation of the reference's own equations (oracle/gsm_oracle_q.c, pinned against mpmath in
    tests/test_oracle_quad.py) on >= 1000 strided targets of the full run: the GPU must hold rtol 1e-10 against it on
    every one of them.  The float64 oracle's own distance to the exact answer is written next to it
    (gpurun_out/exact_c4_uk.json -> profiles/)."""
    import gsm_oracle_c as oc
    o = oracle
    rng = np.random.default_rng(4)
    n = 1_000_000
    X = rng.uniform(0, 256, size=(n, 3))
    z = np.sin(X[:, 0] / 20) + np.cos(X[:, 2] / 30) + 0.1 * rng.standard_normal(n)
    gm = gsm.UniversalKriging(gsm.SphericalVariogram(range=20.0, sill=1.0, nugget=0.1), 2, 3)
    grid = gsm.CartesianGrid(256, 256, 256)
    p1 = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), grid, maxneighbors=32, prob=True).z
    mu1, var1 = np.array(p1.mu), np.array(p1.sigma)
    assert np.isfinite(mu1).all() and (var1 > 0).all()
    om = o.universal_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.1, 20.0), 2, 3)
    idx = np.unique(np.linspace(0, 256 ** 3 - 1, 1200).astype(np.int64))
    assert len(idx) >= 1000
    r = oc.grid_subset_exact(om, X, z, o.Grid((0.0,) * 3, (1.0,) * 3, (256, 256, 256)), idx, 32)
    assert r["ok_q"].all()
    rep = _exact_report("c4_uk", mu1[idx], var1[idx], r)
    assert np.allclose(mu1[idx], r["mu_q"], rtol=RTOL, atol=ATOL), rep
    assert np.allclose(var1[idx], r["var_q"], rtol=RTOL, atol=ATOL), rep


def test_c2_class_gaussian_local_against_exact_arithmetic(gsm, oracle):
    """Gaussian variogram (the n' = nugget + 1e-6 form) with a small nugget, dense samples, local OK k = 32: the GPU
    against the binary128 referee, and how many targets each side flags singular."""
    import gsm_oracle_c as oc
    o = oracle
    X, z = _samples(22, 60_000, 2, 512.0)
    for nug in (0.01, 0.0):
        gm = gsm.OrdinaryKriging(gsm.GaussianVariogram(range=30.0, sill=1.0, nugget=nug))
        grid = gsm.CartesianGrid(256, 256)
        ggrid = gsm.CartesianGrid((0.0, 0.0), (512.0, 512.0), dims=(256, 256))
        pred = gsm.fitpredict(gm, gsm.georef({"z": z}, gsm.PointSet(X)), ggrid, maxneighbors=32, prob=True)
        mu1, var1 = np.array(pred.z.mu), np.array(pred.z.sigma)
        om = o.ordinary_kriging(o.Variogram(o.GAUSSIAN, 1.0, nug, 30.0))
        idx = np.arange(0, 256 * 256, 61)
        r = oc.grid_subset_exact(om, X, z, o.Grid((0.0, 0.0), (2.0, 2.0), (256, 256)), idx, 32)
        gpu_sing = ~np.isfinite(mu1[idx])
        rep = _exact_report("gauss_nug%g" % nug, np.where(gpu_sing, r["mu_q"], mu1[idx]), np.where(gpu_sing, r["var_q"], var1[idx]), r)
        rep["gpu_flagged_singular"] = int(gpu_sing.sum())
        rep["oracle_flagged_singular"] = int((r["st64"] == 2).sum())
        import json, os
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                               "exact_gauss_nug%g.json" % nug), "w") as f:
            json.dump(rep, f, indent=1)
        if nug > 0:
            assert gpu_sing.sum() == 0
            assert np.allclose(mu1[idx], r["mu_q"], rtol=1e-9, atol=1e-11), rep     # kappa ~ 1e4 / nugget
            assert np.allclose(var1[idx], r["var_q"], rtol=1e-9, atol=1e-11), rep
        else:
            # kappa ~ 1e6 .. 1e9 (only the 1e-6 floor regularises): no float64 solver holds 1e-10 here;
            # the bar is the conditioning bound: forward error <= 100 * pivot ratio * 2^-53, on both sides, and the GPU
            # must not flag systems the reference solves
            assert gpu_sing.sum() <= (r["st64"] == 2).sum(), rep
            bound = 100.0 * r["pivot_ratio"].max() * 2.0 ** -53
            e_gpu = np.abs(np.where(gpu_sing, r["mu_q"], mu1[idx]) - r["mu_q"]).max()
            e_o = np.abs(r["mu64"] - r["mu_q"]).max()
            assert e_gpu <= bound and e_o <= bound, rep


def test_randomised_parity_sweep(gsm, oracle):
    """90 random neighbourhood cases (model, variogram, dimension, k up to 64, grid shape, slab, radius,
    minneighbors, coincident samples) against the C oracle: tools/fuzz_parity.py."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location(
        "fuzz_parity", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_parity.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    bad = []
    for seed in range(5000, 5090):
        ok, desc = fz.one(seed)
        if not ok:
            bad.append(desc)
    assert not bad, bad


def test_nn_model_matches_oracle(gsm, oracle):
    """NN (nn.jl:47-54): the nearest sample's value, ties to the lowest row."""
    o = oracle
    rng = np.random.default_rng(12)
    for dim, dims in ((2, (37, 23)), (3, (11, 9, 7))):
        X = np.floor(rng.uniform(0, 30, size=(900, dim)))          # lattice samples: many exact distance ties
        z = rng.standard_normal(900)
        grid = gsm.CartesianGrid(tuple(0.0 for _ in dims), tuple(30.0 for _ in dims), dims=dims)
        og = o.Grid(tuple(0.0 for _ in dims), tuple(30.0 / d for d in dims), dims)
        got = gsm.fitpredict(gsm.NN(), gsm.georef({"z": z}, gsm.PointSet(X)), grid).z
        ref, _, _ = o.fitpredict(o.NNModel(), X, z, og, neighbors=False)
        assert np.array_equal(got, ref)
        P = rng.uniform(-3, 33, size=(500, dim))
        got = gsm.fitpredict(gsm.NN(), gsm.georef({"z": z}, gsm.PointSet(X)), gsm.PointSet(P)).z
        ref, _, _ = o.fitpredict(o.NNModel(), X, z, P, neighbors=False)
        assert np.array_equal(got, ref)


def test_translation_invariance_and_covariance_functions(gsm):
    """test/krig.jl:56-75 (Kriging is translation-invariant) and 314-346 (a covariance function gives the distribution
    of its variogram), through fit / predictprob on the CUDA path, plus the same through fitpredict with neighbours."""
    rng = np.random.default_rng(2024)
    P = rng.uniform(0, 1, size=(100, 3)) * 10.0
    zz = rng.uniform(size=100)
    p0 = np.array([5.0, 5.0, 5.0])
    h = rng.uniform(0, 1, size=3)
    gam, cov = gsm.GaussianVariogram(sill=1.0, range=1.0, nugget=0.0), gsm.GaussianCovariance(sill=1.0, range=1.0, nugget=0.0)
    mk = (lambda f: gsm.SimpleKriging(f, float(zz.mean())), gsm.OrdinaryKriging, lambda f: gsm.UniversalKriging(f, 1, 3),
          lambda f: gsm.UniversalKriging(f, exps=((0, 0, 0),)))
    for make in mk:
        d = gsm.predictprob(gsm.fit(make(gam), gsm.georef({"z": zz}, gsm.PointSet(P))), "z", p0)
        dh = gsm.predictprob(gsm.fit(make(gam), gsm.georef({"z": zz}, gsm.PointSet(P + h))), "z", p0 + h)
        dc = gsm.predictprob(gsm.fit(make(cov), gsm.georef({"z": zz}, gsm.PointSet(P))), "z", p0)
        assert dh.mu == pytest.approx(d.mu, rel=1e-8, abs=1e-10) and dh.sigma == pytest.approx(d.sigma, rel=1e-8)
        assert dc.mu == d.mu and dc.sigma == d.sigma
    X, z = _samples(78, 5000, 2, 100.0)
    fun = gsm.SphericalVariogram(range=30.0, sill=1.2, nugget=0.1)
    grid = gsm.CartesianGrid((0.0, 0.0), (100.0, 100.0), dims=(32, 24))
    shift = np.array([13.25, -7.5])
    gridh = gsm.CartesianGrid(tuple(shift), tuple(shift + 100.0), dims=(32, 24))
    for model in (gsm.OrdinaryKriging(fun), gsm.UniversalKriging(fun, 2, 2)):
        a = gsm.fitpredict(model, gsm.georef({"z": z}, gsm.PointSet(X)), grid, maxneighbors=24, prob=True).z
        b = gsm.fitpredict(model, gsm.georef({"z": z}, gsm.PointSet(X + shift)), gridh, maxneighbors=24, prob=True).z
        assert np.allclose(a.mu, b.mu, rtol=1e-8, atol=1e-10) and np.allclose(a.sigma, b.sigma, rtol=1e-8, atol=1e-10)


def test_fallback_methods_and_cokriging_on_a_grid(gsm):
    """test/krig.jl:431-449 (Symbol / String / tuple forms of predict and predictprob) and the CoKriging block of
    test/misc.jl:27-58 (fitpredict of three variables on a 100 x 100 grid, neighbours on)."""
    d = gsm.georef({"z": np.array([1.0, 0.0, 1.0])}, [(0.0, 0.0), (1.0, 0.0), (1.0, 1.0)])
    ok = gsm.fit(gsm.OrdinaryKriging(gsm.GaussianVariogram()), d)
    p1, p3 = gsm.predict(ok, "z", (0.0, 0.0)), gsm.predict(ok, ("z",), (0.0, 0.0))
    p4, p6 = gsm.predictprob(ok, "z", (0.0, 0.0)), gsm.predictprob(ok, ("z",), (0.0, 0.0))
    assert isinstance(p1, float) and isinstance(p3, list) and p1 == p3[0]
    assert isinstance(p4, gsm.Normal) and isinstance(p6, gsm.MvNormal)
    assert p6.mean()[0] == p4.mu and p6.var()[0] == pytest.approx(p4.sigma ** 2, rel=1e-12)
    B = np.array([[1.0, 0.3, 0.1], [0.3, 1.0, 0.2], [0.1, 0.2, 1.0]])
    model = gsm.Kriging(gsm.CoregionalizedVariogram(B, gsm.SphericalVariogram(range=35.0)))
    data = gsm.georef({"a": np.array([1.0, 0.0, 0.0]), "b": np.array([0.0, 1.0, 0.0]), "c": np.array([0.0, 0.0, 1.0])},
                      [(25.0, 25.0), (50.0, 75.0), (75.0, 50.0)])
    grid = gsm.CartesianGrid((0.5, 0.5), (100.5, 100.5), dims=(100, 100))
    at = {(25, 25): (1.0, 0.0, 0.0), (50, 75): (0.0, 1.0, 0.0), (75, 50): (0.0, 0.0, 1.0)}
    for prob in (False, True):
        pred = gsm.fitpredict(model, data, grid, neighbors=True, prob=prob)
        assert pred.geometry == grid
        for (i, j), want in at.items():
            lin = (i - 1) + (j - 1) * 100          # LinearIndices(size(grid))[i, j], 0-based
            for name, w in zip(("a", "b", "c"), want):
                col = pred.table[name]
                assert abs((col.mu if prob else col)[lin] - w) < 1e-3
