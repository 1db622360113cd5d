"""more variogram families, anisotropic metric balls, nested (composite) structures.
Formulas restated from GeoStatsFunctions 0.15 (see include/gsm_b200.h; unpinned by the reference's own tests, like the
interior of every variogram curve).  CUDA path (general covariance function) against the numpy oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-10, 1e-12


def _rot2(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s], [s, c]])


def _cases(gsm, o):
    """(name, gsm variogram, oracle variogram)"""
    out = []
    for nm, cls, kind in (("cubic", gsm.CubicVariogram, o.CUBIC), ("penta", gsm.PentasphericalVariogram, o.PENTASPHERICAL),
                          ("sinehole", gsm.SineHoleVariogram, o.SINEHOLE), ("circular", gsm.CircularVariogram, o.CIRCULAR)):
        out.append((nm, cls(sill=1.3, range=9.0, nugget=0.15), o.Variogram(kind, 1.3, 0.15, 9.0)))
    R = _rot2(0.6)
    M = np.eye(3)
    M[:2, :2] = np.diag([1 / 12.0, 1 / 4.0]) @ R.T
    out.append(("aniso-spherical", gsm.SphericalVariogram(sill=1.0, nugget=0.1, ranges=(12.0, 4.0), rotation=R),
                o.Variogram(o.SPHERICAL, structs=((o.SPHERICAL, 1.0, 0.1, 1.0, M),))))
    nested = gsm.NuggetEffect(0.1) + 0.6 * gsm.SphericalVariogram(range=5.0) + 0.4 * gsm.GaussianVariogram(range=15.0, nugget=0.05)
    out.append(("nested", nested, o.Variogram(o.SPHERICAL, structs=((o.NUGGET, 0.1, 0.1, 1.0, None),
                                                                    (o.SPHERICAL, 0.6, 0.0, 5.0, None),
                                                                    (o.GAUSSIAN, 0.4, 0.02, 15.0, None)))))
    return out


@pytest.mark.parametrize("idx", range(6))
def test_general_variograms_local_and_global(gsm, oracle, idx):
    o = oracle
    name, gv, ov = _cases(gsm, o)[idx]
    rng = np.random.default_rng(40 + idx)
    X = rng.uniform(0, 1, size=(1200, 2)) * np.array([40.0, 30.0])
    z = np.sin(X[:, 0] / 6) + np.cos(X[:, 1] / 5) + 0.1 * rng.standard_normal(1200)
    tab = gsm.georef({"z": z}, gsm.PointSet(X))
    dims = (21, 17)
    grid = gsm.CartesianGrid((0.0, 0.0), (40.0, 30.0), dims=dims)
    og = o.Grid((0.0, 0.0), (40.0 / dims[0], 30.0 / dims[1]), dims)
    for gm, om, kw in ((gsm.OrdinaryKriging(gv), o.ordinary_kriging(ov), dict(maxneighbors=16)),
                       (gsm.SimpleKriging(gv, 0.2), o.simple_kriging(ov, 0.2), dict(maxneighbors=40)),
                       (gsm.UniversalKriging(gv, 1, 2), o.universal_kriging(ov, 1, 2), dict(maxneighbors=24)),
                       (gsm.OrdinaryKriging(gv), o.ordinary_kriging(ov), dict(neighbors=False))):
        sub = X[:300], z[:300]
        Xs, zs = (sub if kw.get("neighbors") is False else (X, z))
        t = gsm.georef({"z": zs}, gsm.PointSet(Xs))
        mu, var, st = o.fitpredict(om, Xs, zs, og, prob=True, **kw)
        got = gsm.fitpredict(gm, t, grid, prob=True, **kw).z
        assert (st == 0).all(), name
        tol = RTOL if kw.get("neighbors", True) else 1e-8
        assert np.allclose(got.mu, mu, rtol=tol, atol=tol * 1e-2), (name, kw, np.abs(got.mu - mu).max())
        assert np.allclose(got.sigma, var, rtol=tol, atol=tol * 1e-2), (name, kw, np.abs(got.sigma - var).max())


def test_anisotropic_3d_nested(gsm, oracle):
    o = oracle
    rng = np.random.default_rng(77)
    X = rng.uniform(0, 20, size=(4000, 3))
    z = np.sin(X[:, 0] / 4) + 0.2 * X[:, 2] / 20 + 0.1 * rng.standard_normal(4000)
    # rotation about z by 30 degrees, radii (8, 4, 2); plus an isotropic exponential structure
    R = np.eye(3)
    R[:2, :2] = _rot2(np.pi / 6)
    gv = 0.7 * gsm.SphericalVariogram(nugget=0.05, ranges=(8.0, 4.0, 2.0), rotation=R) + 0.3 * gsm.ExponentialVariogram(range=6.0)
    M = np.diag([1 / 8.0, 1 / 4.0, 1 / 2.0]) @ R.T
    ov = o.Variogram(o.SPHERICAL, structs=((o.SPHERICAL, 0.7, 0.035, 1.0, M), (o.EXPONENTIAL, 0.3, 0.0, 6.0, None)))
    dims = (9, 8, 7)
    grid = gsm.CartesianGrid((0.0,) * 3, (20.0,) * 3, dims=dims)
    og = o.Grid((0.0,) * 3, tuple(20.0 / d for d in dims), dims)
    mu, var, st = o.fitpredict(o.ordinary_kriging(ov), X, z, og, prob=True, maxneighbors=32)
    got = gsm.fitpredict(gsm.OrdinaryKriging(gv), gsm.georef({"z": z}, gsm.PointSet(X)), grid, maxneighbors=32, prob=True).z
    assert (st == 0).all()
    assert np.allclose(got.mu, mu, rtol=RTOL, atol=ATOL) and np.allclose(got.sigma, var, rtol=RTOL, atol=ATOL)
