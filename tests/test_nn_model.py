"""The NN model (src/nn.jl) through fit / predict / predictprob and fitpredict: the reference's "Basics" testset
(test/nn.jl:2-31: string values on the default 1-D grid of georef, `missing` values skipped) restated on the oracle (CPU)
and run on the CUDA path (values cross the C ABI as category codes; samples without a value are not uploaded)."""
import numpy as np
import pytest


def test_oracle_nn_skips_missing(oracle):
    X = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]])
    z = np.array([np.nan, np.nan, 7.0])
    for p in X:                                   # test/nn.jl:22-30
        assert oracle.nn_predict(X, z, p) == 7.0
    assert np.isnan(oracle.nn_predict(X, np.full(3, np.nan), X[0]))
    assert oracle.nn_predict(X, np.array([1.0, 2.0, 3.0]), np.array([0.4, 0.0])) == 1.0


def test_value_encoding_and_default_domain(gsm):
    """host logic: category codes / missing markers, georef without a domain, centroids of grid data"""
    api = __import__("gsm_b200.api", fromlist=["_encode"])
    z, cats = api._encode(["b", None, "a", "b"])
    assert list(cats) == ["a", "b"] and np.array_equal(z, [1.0, np.nan, 0.0, 1.0], equal_nan=True)
    z, cats = api._encode([1, None, 2.5])
    assert cats is None and np.array_equal(z, [1.0, np.nan, 2.5], equal_nan=True)
    z, cats = api._encode(np.ma.masked_array([1.0, 2.0], mask=[False, True]))
    assert cats is None and z[0] == 1.0 and np.isnan(z[1])
    d = gsm.georef({"z": ["a", "b", "c"]})
    assert d.domain == gsm.CartesianGrid(3)
    assert np.array_equal(api._pointset_of(d.domain, centroids_ok=True), [[0.5], [1.5], [2.5]])
    g = gsm.CartesianGrid((0.0, 10.0), (4.0, 16.0), dims=(2, 3))
    assert np.array_equal(api._pointset_of(g, centroids_ok=True),
                          [[1.0, 11.0], [3.0, 11.0], [1.0, 13.0], [3.0, 13.0], [1.0, 15.0], [3.0, 15.0]])
    with pytest.raises(NotImplementedError):
        api._pointset_of(g)


@pytest.mark.gpu
def test_nn_basics_reference_testset(gsm):
    data = gsm.georef({"z": ["a", "b", "c"]})                       # test/nn.jl:3-10: cells of CartesianGrid(3)
    nn = gsm.fit(gsm.NN(), data)
    assert gsm.status(nn)
    for x, want in ((0.5, "a"), (1.5, "b"), (2.5, "c"), (0.9, "a"), (1.1, "b")):
        assert gsm.predict(nn, "z", (x,)) == want
        assert gsm.predictprob(nn, "z", (x,)).value == want
    data = gsm.georef({"z": [None, None, "c"]}, [(0.0, 0.0), (1.0, 0.0), (2.0, 0.0)])   # test/nn.jl:22-30
    nn = gsm.fit(gsm.NN(), data)
    for p in ((0.0, 0.0), (1.0, 0.0), (2.0, 0.0)):
        assert gsm.predict(nn, "z", p) == "c"
    nn = gsm.fit(gsm.NN(), gsm.georef({"z": [None, None]}, [(0.0, 0.0), (1.0, 0.0)]))
    assert gsm.predict(nn, "z", (0.0, 0.0)) is None                # `missing`
    nn = gsm.fit(gsm.NN(), gsm.georef({"z": [1.5, np.nan, -2.0]}, [(0.0, 0.0), (1.0, 0.0), (2.0, 0.0)]))
    assert gsm.predict(nn, "z", (0.9, 0.0)) == 1.5 and gsm.predict(nn, ["z"], (1.6, 0.0)) == [-2.0]


@pytest.mark.gpu
def test_nn_fitpredict_categories_missing_and_grid_data(gsm, oracle):
    o = oracle
    rng = np.random.default_rng(12)
    X = rng.uniform(0, 50, size=(3000, 2))
    lab = np.array(["sand", "clay", "silt", "rock"], dtype=object)[rng.integers(0, 4, 3000)]
    lab[rng.random(3000) < 0.2] = None
    grid = gsm.CartesianGrid((0.0, 0.0), (50.0, 50.0), dims=(60, 40))
    og = o.Grid((0.0, 0.0), (50.0 / 60, 50.0 / 40), (60, 40))
    pred = gsm.fitpredict(gsm.NN(), gsm.georef({"z": list(lab)}, gsm.PointSet(X)), grid, neighbors=False).z
    cats = sorted({v for v in lab if v is not None})
    codes = np.array([np.nan if v is None else cats.index(v) for v in lab])
    ref, _, _ = o.fitpredict(o.NNModel(), X, codes, og, neighbors=False)
    assert [cats[int(c)] for c in ref] == list(pred)
    # numeric values on grid data (centroids, models.jl:265-270), default neighbourhood search
    dgrid = gsm.CartesianGrid((0.0, 0.0), (50.0, 50.0), dims=(25, 20))
    z = rng.standard_normal(500)
    Xc = o.Grid((0.0, 0.0), (2.0, 2.5), (25, 20)).centroids()
    for m, om in ((gsm.NN(), o.NNModel()), (gsm.IDW(2), o.IDWModel(2))):
        got = gsm.fitpredict(m, gsm.georef({"z": z}, dgrid), grid, maxneighbors=6).z
        ref, _, _ = o.fitpredict(om, Xc, z, og, maxneighbors=6)
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-13)
    ok = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=10.0, nugget=0.05))
    got = gsm.fitpredict(ok, gsm.georef({"z": z}, dgrid), grid, maxneighbors=12).z
    ref, _, _ = o.fitpredict(o.ordinary_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.05, 10.0)), Xc, z, og, maxneighbors=12)
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-12)
    with pytest.raises(NotImplementedError):          # data with volume support under point=false: not claimed
        gsm.fitpredict(ok, gsm.georef({"z": z}, dgrid), grid, maxneighbors=12, point=False)


@pytest.mark.gpu
def test_nn_with_missing_values_behind_a_neighbour_search(gsm, oracle):
    """fitpredictneigh with NN over a column with `missing` (models.jl:162-184, nn.jl:47-54): nearest of the k nearest
    that has a value; `missing` when none has, or when the ball holds fewer than minneighbors samples."""
    o = oracle
    rng = np.random.default_rng(23)
    X = rng.uniform(0, 40, size=(4000, 2))
    z = np.round(rng.standard_normal(4000), 3)
    z[rng.random(4000) < 0.6] = np.nan                     # most samples have no value
    grid = gsm.CartesianGrid((0.0, 0.0), (40.0, 40.0), dims=(50, 40))
    og = o.Grid((0.0, 0.0), (0.8, 1.0), (50, 40))
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    for kw in (dict(maxneighbors=3), dict(maxneighbors=10), dict(maxneighbors=6, minneighbors=4, neighborhood=1.2)):
        got = gsm.fitpredict(gsm.NN(), tab, grid, **kw).z
        ref, _, st = o.fitpredict(o.NNModel(), X, z, og, **kw)
        gone = np.isnan(ref) | (st == 1)
        assert 0 < gone.sum() < len(ref), kw
        assert np.array_equal(np.ma.getmaskarray(got), gone), kw
        assert np.array_equal(np.ma.getdata(got)[~gone], ref[~gone]), kw
    lab = np.array([None if np.isnan(v) else ("hi" if v > 0 else "lo") for v in z], dtype=object)
    got = gsm.fitpredict(gsm.NN(), gsm.georef({"z": list(lab)}, gsm.PointSet(X)), grid, maxneighbors=3).z
    ref, _, _ = o.fitpredict(o.NNModel(), X, np.where(np.isnan(z), np.nan, (z > 0) * 1.0), og, maxneighbors=3)
    assert np.array_equal(np.ma.getmaskarray(got), np.isnan(ref))
    assert list(np.ma.getdata(got)[~np.isnan(ref)]) == [("hi" if c == 1.0 else "lo") for c in ref[~np.isnan(ref)]]
