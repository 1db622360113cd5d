"""The multi-GPU entry points of the C ABI (gsm_multi_*): one caller, all GPUs of the box, samples replicated by one
NCCL broadcast, target slabs.  Runs on whatever the box has: with one GPU the handle degenerates to one slab (no
NCCL); with >= 2 GPUs the slabs must reproduce the single-GPU result and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("ndev", [1, 2, 4, 8])
def test_multi_local_kriging_matches_single_device_and_oracle(gsm, oracle, ndev):
    if _ndev() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    import gsm_oracle_c as oc
    o = oracle
    rng = np.random.default_rng(91)
    for dim, dims, n in ((2, (640, 512), 40_000), (3, (64, 48, 40), 60_000)):
        ext = np.array(dims, dtype=float)
        X = rng.uniform(0, 1, size=(n, dim)) * ext
        z = np.sin(X[:, 0] / 20) + np.cos(X[:, -1] / 30) + 0.1 * rng.standard_normal(n)
        gm = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=25.0, sill=1.0, nugget=0.1))
        tab = gsm.georef({"z": z}, gsm.PointSet(X))
        grid = gsm.CartesianGrid(*dims)
        one = gsm.fitpredict(gm, tab, grid, maxneighbors=32, prob=True).z
        many = gsm.fitpredict(gm, tab, grid, maxneighbors=32, prob=True, devices=list(range(ndev))).z
        # slabs are aligned to patch boundaries, so every system sees the same patch as in the single-device run
        assert np.array_equal(np.asarray(one.mu), np.asarray(many.mu))
        assert np.array_equal(np.asarray(one.sigma), np.asarray(many.sigma))
        om = o.ordinary_kriging(o.Variogram(o.SPHERICAL, 1.0, 0.1, 25.0))
        M = int(np.prod(dims))
        first, count = (M // 3) // dims[0] * dims[0], 5000
        mu, var, st, _, _ = oc.fitpredict(om, X, z, o.Grid(tuple(0.0 for _ in dims), tuple(1.0 for _ in dims), dims),
                                          prob=True, maxneighbors=32, first=first, count=count, nthreads=8)
        assert np.allclose(np.asarray(many.mu)[first:first + count], mu, rtol=1e-10, atol=1e-12)
        assert np.allclose(np.asarray(many.sigma)[first:first + count], var, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_slabs_partition_idw_global_and_shards(gsm, oracle, ndev):
    if _ndev() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    L = gsm._lib
    mc = gsm.multi_context(list(range(ndev)))
    rng = np.random.default_rng(92)
    X = rng.uniform(0, 100, size=(3000, 2))
    z = rng.standard_normal(3000)
    # the slabs of any target range are contiguous, disjoint and cover it
    for dims, first, count in (((333, 77), 0, -1), ((333, 77), 1234, 9999), ((50, 40, 30), 0, -1), ((7, 5), 3, 20)):
        t = L.Context.grid_targets(tuple(0.0 for _ in dims), tuple(1.0 for _ in dims), dims, first, count)
        total = L.Context.target_count(t)
        pos = first
        for i in range(ndev):
            f, c = mc.slab(t, i)
            assert f == pos and c >= 0
            pos += c
        assert pos == first + total
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    grid = gsm.CartesianGrid((0.0, 0.0), (100.0, 100.0), dims=(120, 90))
    for model, kw in ((gsm.IDW(2.0), dict(maxneighbors=12)), (gsm.IDW(2.0), dict(neighbors=False)),
                      (gsm.OrdinaryKriging(gsm.GaussianVariogram(range=15.0, nugget=0.05)), dict(neighbors=False, prob=True)),
                      (gsm.SimpleKriging(gsm.ExponentialVariogram(range=20.0, nugget=0.05), 0.1), dict(maxneighbors=40, prob=True))):
        a = gsm.fitpredict(model, tab, grid, **kw).z
        b = gsm.fitpredict(model, tab, grid, devices=list(range(ndev)), **kw).z
        if kw.get("prob"):
            assert np.allclose(a.mu, b.mu, rtol=1e-12, atol=1e-13) and np.allclose(a.sigma, b.sigma, rtol=1e-12, atol=1e-13)
        else:
            assert np.allclose(np.asarray(a), np.asarray(b), rtol=1e-12, atol=1e-13)
    if ndev > 1:     # (a one-device list goes through the plain context)
        t = mc.timings()
        assert t["wall_ms"] > 0 and t["predict_ms"] > 0


def test_multi_and_option_error_paths(gsm):
    """Error behaviour of the round-2 entry points: codes, messages, no exceptions across the C ABI."""
    import ctypes as C
    L = gsm._lib
    lib = L.load()
    h = C.c_void_p()
    # duplicate / out-of-range device ordinals
    devs = (C.c_int32 * 2)(0, 0)
    assert lib.gsm_multi_create(devs, 2, C.byref(h)) == L.GSM_ERR_INVALID and b"duplicate" in lib.gsm_multi_last_error(None)
    devs = (C.c_int32 * 1)(99)
    assert lib.gsm_multi_create(devs, 1, C.byref(h)) == L.GSM_ERR_INVALID
    assert lib.gsm_multi_create(devs, 0, C.byref(h)) == L.GSM_ERR_INVALID
    # predict before set_samples
    mc = L.MultiContext([0])
    t = L.Context.grid_targets((0.0, 0.0), (1.0, 1.0), (8, 8))
    with pytest.raises(gsm.GsmError) as e:
        mc.local_krige(L.make_variogram(1, 1, 0, 1), L.make_model(L.GSM_KRIG_ORDINARY), t, 8)
    assert e.value.code == L.GSM_ERR_NO_SAMPLES
    mc.close()
    # options: unknown key / value rejected, known ones accepted and effective
    c = gsm.Context(0)
    with pytest.raises(gsm.GsmError):
        c.set_option("no_such_option", 1)
    with pytest.raises(gsm.GsmError):
        c.set_option("local_impl", "fastest")
    rng = np.random.default_rng(1)
    X = rng.uniform(0, 40, size=(3000, 2))
    z = rng.standard_normal(3000)
    model = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=8.0, nugget=0.1))
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    grid = gsm.CartesianGrid((0.0, 0.0), (40.0, 40.0), dims=(96, 80))
    ref = gsm.fitpredict(model, tab, grid, maxneighbors=32, prob=True, ctx=c).z
    for impl in ("patch", "patch2", "pipe", "simt", "shr", "auto"):
        c.set_option("local_impl", impl)
        got = gsm.fitpredict(model, tab, grid, maxneighbors=32, prob=True, ctx=c).z
        assert np.allclose(got.mu, ref.mu, rtol=1e-11, atol=1e-12) and np.allclose(got.sigma, ref.sigma, rtol=1e-11, atol=1e-12), impl
    c.close()


@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_change_of_support(gsm, ndev):
    """point=false through the multi-GPU handle: the block table goes to every device's context."""
    if _ndev() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    rng = np.random.default_rng(93)
    dims = (96, 64)
    X = rng.uniform(0, 1, size=(5000, 2)) * np.array(dims, dtype=float)
    z = np.sin(X[:, 0] / 9) + 0.1 * rng.standard_normal(5000)
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    grid = gsm.CartesianGrid(*dims)
    gm = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=12.0, nugget=0.1))
    for kw in (dict(maxneighbors=16), dict(neighbors=False)):
        data = tab if kw.get("neighbors", True) else gsm.georef({"z": z[:600]}, gsm.PointSet(X[:600]))
        one = gsm.fitpredict(gm, data, grid, point=False, prob=True, **kw).z
        many = gsm.fitpredict(gm, data, grid, point=False, prob=True, devices=list(range(ndev)), **kw).z
        pt = gsm.fitpredict(gm, data, grid, point=True, prob=True, devices=list(range(ndev)), **kw).z
        assert np.array_equal(np.asarray(one.mu), np.asarray(many.mu))
        assert np.array_equal(np.asarray(one.sigma), np.asarray(many.sigma))
        assert np.abs(np.asarray(pt.sigma) - np.asarray(many.sigma)).max() > 1e-6   # (and the table was cleared again)
