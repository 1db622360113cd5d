"""Host-side mirror of the GeoStatsModels.jl interface for the interpolation hot path.

Same names, argument meaning and error behaviour as the reference (paths under /root/reference/src):
``fit`` (models.jl:38), ``predict`` (models.jl:54-61), ``predictprob`` (models.jl:69-76), ``status``
(models.jl:84), ``fitpredict`` (models.jl:102-129) and the models ``IDW`` (idw.jl:20-26),
``SimpleKriging`` (krig/simple.jl:29-45), ``OrdinaryKriging`` (krig/ordinary.jl:25-27),
``UniversalKriging`` (krig/universal.jl:45-50), ``Kriging`` (krig.jl:423-427) with
Gaussian/Spherical/Exponential variograms.  Julia is not installed in this image, so this Python layer
stands where the Julia shim of INTEGRATION.md would: it does the orchestration of models.jl (point
support, `_scale`, neighbour-limit clamping) and hands whole target domains to the C ABI.
Everything numerical happens in libgsm_b200.so on the GPU; there is no CPU path here.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass

import numpy as np

from . import _lib

__all__ = [
    "CoregionalizedVariogram", "GaussianVariogram", "SphericalVariogram", "ExponentialVariogram", "CubicVariogram", "PentasphericalVariogram",
    "SineHoleVariogram", "CircularVariogram", "NuggetEffect", "CompositeVariogram", "PointSet", "CartesianGrid", "GeoTable",
    "georef", "IDW", "NN", "SimpleKriging", "OrdinaryKriging", "UniversalKriging", "Kriging", "Normal", "fit",
    "predict", "predictprob", "status", "fitpredict", "default_context", "multi_context", "set_default_device", "Dirac",
    "Box", "Quadrangle", "block_points", "MvNormal",
    "GaussianCovariance", "SphericalCovariance", "ExponentialCovariance", "CubicCovariance", "PentasphericalCovariance",
    "SineHoleCovariance", "CircularCovariance",
]


# ---------------------------------------------------------------------------------------------
# geostatistical functions (the GeoStatsFunctions subset the GPU path claims)
# ---------------------------------------------------------------------------------------------
class _Variogram:
    """One GeoStatsFunctions variogram structure: family, total sill, nugget and a metric ball -- isotropic
    (`range=r`) or anisotropic (`ranges=(r1, r2[, r3])` with an optional rotation matrix `rotation`, the columns being
    the ball's principal axes: MetricBall(ranges, rotation))."""
    kind = -1

    def __init__(self, sill=1.0, range=1.0, nugget=0.0, ranges=None, rotation=None):
        self.sill, self.nugget = float(sill), float(nugget)
        self.ranges = None if ranges is None else tuple(float(r) for r in ranges)
        self.rotation = None if rotation is None else np.asarray(rotation, dtype=np.float64)
        self.range = float(max(self.ranges)) if self.ranges is not None else float(range)   # range(fun): largest radius

    def __eq__(self, other):
        return type(self) is type(other) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def _key(self):
        return (self.kind, self.sill, self.nugget, self.range, self.ranges,
                None if self.rotation is None else tuple(self.rotation.ravel()))

    def scaled(self, alpha):
        """GeoStatsFunctions.scale(fun, alpha): the radii of the metric ball are multiplied by alpha."""
        if self.ranges is None:
            return type(self)(sill=self.sill, range=alpha * self.range, nugget=self.nugget)
        return type(self)(sill=self.sill, nugget=self.nugget, ranges=tuple(alpha * r for r in self.ranges),
                          rotation=self.rotation)

    def _structure(self, coef=1.0):
        """(kind, sill, nugget, range, metric) for the C descriptor; coef scales sill and nugget (c * gamma)."""
        if self.ranges is None:
            return (self.kind, coef * self.sill, coef * self.nugget, self.range, None)
        d = len(self.ranges)
        R = np.eye(3)
        if self.rotation is not None:
            R[:d, :d] = self.rotation[:d, :d]
        inv = np.zeros(3)
        inv[:d] = 1.0 / np.asarray(self.ranges)
        return (self.kind, coef * self.sill, coef * self.nugget, 1.0, np.diag(inv) @ R.T)

    def _structures(self):
        return [self._structure()]

    def _c(self):
        if self.ranges is None and self.kind <= _lib.GSM_VARIO_NUGGET:
            return _lib.make_variogram(self.kind, self.sill, self.nugget, self.range)
        return _lib.make_composite_variogram(self._structures())

    # c * gamma and gamma1 + gamma2 (GeoStatsFunctions CompositeFunction: a sum of scaled structures)
    def __rmul__(self, c):
        return CompositeVariogram([(float(c), self)])

    def __add__(self, other):
        return CompositeVariogram([(1.0, self)]) + other


class CompositeVariogram:
    """Nested variogram sum_i c_i gamma_i (every structure with its own family, sill, nugget and metric ball)."""

    def __init__(self, terms):
        self.terms = [(float(c), f) for c, f in terms]
        if not 1 <= len(self.terms) <= _lib.GSM_MAX_STRUCTURES:
            raise ValueError("a composite variogram takes 1..4 structures")

    @property
    def range(self):       # range(fun): the largest radius over the structures
        return max(f.range for _, f in self.terms)

    @property
    def sill(self):
        return sum(c * f.sill for c, f in self.terms)

    def scaled(self, alpha):
        return CompositeVariogram([(c, f.scaled(alpha)) for c, f in self.terms])

    def _structures(self):
        return [f._structure(c) for c, f in self.terms]

    def _c(self):
        return _lib.make_composite_variogram(self._structures())

    def __add__(self, other):
        more = other.terms if isinstance(other, CompositeVariogram) else [(1.0, other)]
        return CompositeVariogram(self.terms + list(more))

    def __rmul__(self, c):
        return CompositeVariogram([(float(c) * a, f) for a, f in self.terms])


class CoregionalizedVariogram:
    """B * gamma0: a matrix coefficient times a scalar variogram (GeoStatsFunctions: `[1.0 0.3; 0.3 1.0] * gamma`), the
    multivariate function of the reference's CoKriging tests (test/krig.jl:348-395).  With every variable observed at
    every sample the cokriging system (C0 (x) B) decouples: the weights of every variable are the scalar Kriging
    weights of gamma0 and the covariance of the prediction is B times the scalar variance ("autokrigeability" of the
    proportional model; tests/test_cokriging.py checks it against the reference's full (nobs nvar + ncon nvar) system
    restated by the CPU test oracle).  The GPU path therefore serves this form with one univariate solve per variable."""

    def __init__(self, B, base):
        self.B = np.asarray(B, dtype=np.float64)
        if self.B.ndim != 2 or self.B.shape[0] != self.B.shape[1] or not np.allclose(self.B, self.B.T):
            raise ValueError("the coefficient must be a symmetric matrix")
        self.base = base

    @property
    def nvariables(self):
        return self.B.shape[0]

    @property
    def range(self):
        return self.base.range

    def scaled(self, alpha):
        return CoregionalizedVariogram(self.B, self.base.scaled(alpha))

    def marginal(self, p):
        """the variogram of variable p alone: B[p, p] * gamma0"""
        return float(self.B[p, p]) * self.base


class GaussianVariogram(_Variogram):
    kind = _lib.GSM_VARIO_GAUSSIAN


class SphericalVariogram(_Variogram):
    kind = _lib.GSM_VARIO_SPHERICAL


class ExponentialVariogram(_Variogram):
    kind = _lib.GSM_VARIO_EXPONENTIAL


class CubicVariogram(_Variogram):
    kind = _lib.GSM_VARIO_CUBIC


class PentasphericalVariogram(_Variogram):
    kind = _lib.GSM_VARIO_PENTASPHERICAL


class SineHoleVariogram(_Variogram):
    kind = _lib.GSM_VARIO_SINEHOLE


class CircularVariogram(_Variogram):
    kind = _lib.GSM_VARIO_CIRCULAR


class NuggetEffect(_Variogram):
    """NuggetEffect(nugget): gamma(h) = (h > 0) nugget"""
    kind = _lib.GSM_VARIO_NUGGET

    def __init__(self, nugget=1.0, **kw):
        super().__init__(sill=nugget, range=1.0, nugget=nugget)

    def scaled(self, alpha):
        return NuggetEffect(self.nugget)


# Covariance functions cov(h) = sill - gamma(h) of the same families (GeoStatsFunctions `GaussianCovariance`, ...).  The
# reference skips lhsbanded! / rhsbanded! for them (`isbanded`, krig.jl:165-179, 364-378) because they are already in the
# covariance form the GPU path always works in, so they cross the C ABI as the descriptor of their variogram
# (test/krig.jl:314-346: "distributions match no matter the geostats function").
class GaussianCovariance(GaussianVariogram):
    pass


class SphericalCovariance(SphericalVariogram):
    pass


class ExponentialCovariance(ExponentialVariogram):
    pass


class CubicCovariance(CubicVariogram):
    pass


class PentasphericalCovariance(PentasphericalVariogram):
    pass


class SineHoleCovariance(SineHoleVariogram):
    pass


class CircularCovariance(CircularVariogram):
    pass


# ---------------------------------------------------------------------------------------------
# domains and geotables (the Meshes / GeoTables subset the GPU path claims)
# ---------------------------------------------------------------------------------------------
class PointSet:
    def __init__(self, coords):
        c = np.asarray(coords, dtype=np.float64)
        if c.ndim == 1:
            c = c[:, None]
        if c.ndim != 2 or not 1 <= c.shape[1] <= 3:
            raise ValueError("PointSet needs an (n, dim) coordinate array with dim in 1..3")
        self.coords = np.ascontiguousarray(c)

    @property
    def dim(self):
        return self.coords.shape[1]

    def nelements(self):
        return self.coords.shape[0]

    def __eq__(self, other):
        return isinstance(other, PointSet) and np.array_equal(self.coords, other.coords)

    def _absmax(self):  # _scalefactor(domain), models.jl:286-291
        return float(np.abs(self.coords).max())


class CartesianGrid:
    """CartesianGrid(nx, ny[, nz]) -> origin 0, unit spacing;
    CartesianGrid(lower, upper, dims=(..)) -> spacing (upper-lower)/dims, as in Meshes."""

    def __init__(self, *args, dims=None):
        if dims is None:
            self.dims = tuple(int(a) for a in args)
            self.origin = tuple(0.0 for _ in self.dims)
            self.spacing = tuple(1.0 for _ in self.dims)
        else:
            lo, hi = args
            self.dims = tuple(int(d) for d in dims)
            self.origin = tuple(float(x) for x in lo)
            self.spacing = tuple((float(h) - float(l)) / d for l, h, d in zip(lo, hi, self.dims))
        if not 1 <= len(self.dims) <= 3:
            raise ValueError("CartesianGrid supports 1 to 3 dimensions")

    @property
    def dim(self):
        return len(self.dims)

    def nelements(self):
        return int(np.prod(self.dims))

    def __eq__(self, other):
        return (isinstance(other, CartesianGrid) and self.dims == other.dims and self.origin == other.origin
                and self.spacing == other.spacing)

    def _absmax(self):  # bounding box corners, models.jl:286-291
        lo = np.asarray(self.origin)
        hi = lo + np.asarray(self.dims, dtype=np.float64) * np.asarray(self.spacing)
        return float(max(np.abs(lo).max(), np.abs(hi).max()))


class Box:
    """Box(lower, upper): an axis-aligned segment / quadrangle / hexahedron as a prediction target with volume support
    (the `Quadrangle((0,0),(1,0),(1,1),(0,1))` of test/krig.jl:231 is Box((0, 0), (1, 1)))."""

    def __init__(self, lower, upper):
        self.lower = tuple(float(x) for x in lower)
        self.upper = tuple(float(x) for x in upper)
        if len(self.lower) != len(self.upper) or not 1 <= len(self.lower) <= 3:
            raise ValueError("Box needs two corners of the same dimension (1 to 3)")

    @property
    def sides(self):
        return tuple(h - l for l, h in zip(self.lower, self.upper))

    def centroid(self):
        return np.array([(l + h) / 2 for l, h in zip(self.lower, self.upper)])


def Quadrangle(a, b, c, d) -> "Box":
    """Quadrangle(a, b, c, d) with axis-aligned sides, as a Box (other quadrangles are not claimed)."""
    P = np.asarray([a, b, c, d], dtype=np.float64)
    lo, hi = P.min(axis=0), P.max(axis=0)
    corners = {(x, y) for x in (lo[0], hi[0]) for y in (lo[1], hi[1])}
    if P.shape != (4, 2) or {tuple(p_) for p_ in P} != corners:
        raise NotImplementedError("only axis-aligned quadrangles are claimed by the GPU path")
    return Box(lo, hi)


class GeoTable:
    """georef(table, domain): named columns of values attached to a domain."""

    def __init__(self, table: dict, domain):
        self.table = {k: v for k, v in table.items()}
        self.domain = domain
        for k, v in self.table.items():
            if len(v) != domain.nelements():
                raise ValueError(f"column {k!r} does not have one value per element")

    @property
    def geometry(self):
        return self.domain

    def __getattr__(self, name):
        t = self.__dict__.get("table", {})
        if name in t:
            return t[name]
        raise AttributeError(name)

    def names(self):
        return tuple(self.table.keys())

    def __eq__(self, other):
        return (isinstance(other, GeoTable) and self.domain == other.domain and self.names() == other.names()
                and all(_col_equal(self.table[k], other.table[k]) for k in self.table))


def _col_equal(a, b):
    if isinstance(a, Normal) or isinstance(b, Normal):
        return isinstance(a, Normal) and isinstance(b, Normal) and np.array_equal(a.mu, b.mu) and np.array_equal(a.sigma, b.sigma)
    return np.array_equal(np.ma.filled(a, np.nan), np.ma.filled(b, np.nan), equal_nan=True)


def georef(table: dict, domain=None) -> GeoTable:
    if domain is None:      # georef of plain vectors: a 1-D CartesianGrid with one cell per row (GeoTables)
        domain = CartesianGrid(len(next(iter(table.values()))))
    if not isinstance(domain, (PointSet, CartesianGrid)):
        domain = PointSet(domain)
    return GeoTable(table, domain)


@dataclass
class Normal:
    """Normal(mu, sigma), scalar or column.  NOTE (reference quirk, copied not fixed): on the
    ``fitpredict(prob=true)`` path models.jl:298-299 builds ``Normal.(mean, var)``, so ``sigma`` there
    holds the Kriging VARIANCE (+1e-10); ``predictprob(fitted, :z, p)`` (krig.jl:223-229) holds sqrt."""
    mu: object
    sigma: object

    def mean(self):
        return self.mu


class MvNormal:
    """MvNormal(mu, Sigma): the joint predictive distribution `predictprob(fitted, vars, g)` returns for several
    variables (krig.jl:231-237); `var()` is the diagonal, as Distributions.var sees it."""

    def __init__(self, mu, Sigma):
        self.mu = np.asarray(mu, dtype=np.float64)
        self.Sigma = np.asarray(Sigma, dtype=np.float64)
        np.linalg.cholesky(self.Sigma)           # PosDefException of MvNormal(mu, Sigma), krig.jl:236

    def mean(self):
        return self.mu

    def var(self):
        return np.diag(self.Sigma).copy()

    def cov(self):
        return self.Sigma

    def __repr__(self):
        return f"MvNormal(mu={self.mu!r}, Sigma={self.Sigma!r})"


# ---------------------------------------------------------------------------------------------
# models
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class IDW:
    exponent: float = 1
    distance: str = "euclidean"

    def _range(self):
        return math.inf  # Base.range(::GeoStatsModel) = Inf (models.jl:22)


@dataclass(frozen=True)
class NN:
    """NN(distance=Euclidean()) (nn.jl:14-24): the value of the nearest sample."""
    distance: str = "euclidean"

    def _range(self):
        return math.inf



class _KrigingModel:
    fun: _Variogram

    def _range(self):  # krig.jl:12
        return self.fun.range


@dataclass(frozen=True)
class SimpleKriging(_KrigingModel):
    fun: _Variogram
    mean: float

    def _scaled(self, a):
        return SimpleKriging(self.fun.scaled(a), self.mean)

    def _c(self):
        return _lib.make_model(_lib.GSM_KRIG_SIMPLE, mean=self.mean)


@dataclass(frozen=True)
class OrdinaryKriging(_KrigingModel):
    fun: _Variogram

    def _scaled(self, a):
        return OrdinaryKriging(self.fun.scaled(a))

    def _c(self):
        return _lib.make_model(_lib.GSM_KRIG_ORDINARY)


def monomial_exponents(deg: int, dim: int):
    """universal.jl:68-77: multiexponents per degree, stably sorted by maximum exponent (descending)."""
    assert deg >= 0, "degree must be nonnegative"   # universal.jl:56
    assert dim > 0, "dimension must be positive"     # universal.jl:57
    cols = []
    for d in range(deg + 1):
        cols.extend(e for e in itertools.product(range(d, -1, -1), repeat=dim) if sum(e) == d)
    order = sorted(range(len(cols)), key=lambda i: -max(cols[i]))
    return tuple(cols[i] for i in order)


class UniversalKriging(_KrigingModel):
    """UniversalKriging(fun, deg, dim) (universal.jl:50).  Arbitrary drift closures (universal.jl:26)
    cannot cross a C ABI; pass a table of monomial exponents instead: UniversalKriging(fun, exps=[...])."""

    def __init__(self, fun, deg=None, dim=None, exps=None):
        self.fun = fun
        if exps is None:
            exps = monomial_exponents(int(deg), int(dim))
        self.exps = tuple(tuple(int(x) for x in e) for e in exps)
        if not 1 <= len(self.exps) <= _lib.GSM_MAX_DRIFT:
            raise ValueError("number of drift monomials must be in 1..10")

    def _scaled(self, a):
        return UniversalKriging(self.fun.scaled(a), exps=self.exps)

    def _c(self):
        return _lib.make_model(_lib.GSM_KRIG_UNIVERSAL, exps=self.exps)


def Kriging(fun, *args):
    """krig.jl:423-427: Kriging(f) -> OK, Kriging(f, mu) -> SK, Kriging(f, deg, dim) -> UK."""
    if len(args) == 0:
        return OrdinaryKriging(fun)
    if len(args) == 1 and isinstance(args[0], (int, float)) and not isinstance(args[0], bool):
        return SimpleKriging(fun, float(args[0]))
    if len(args) == 2:
        return UniversalKriging(fun, int(args[0]), int(args[1]))
    if len(args) == 1:
        return UniversalKriging(fun, exps=args[0])
    raise TypeError("unsupported Kriging constructor arguments")


# ---------------------------------------------------------------------------------------------
# device context management (one context per GPU per process)
# ---------------------------------------------------------------------------------------------
_CTX: dict[int, _lib.Context] = {}
_DEFAULT_DEVICE = 0


def set_default_device(device: int):
    global _DEFAULT_DEVICE
    _DEFAULT_DEVICE = int(device)


def default_context(device: int | None = None) -> _lib.Context:
    dev = _DEFAULT_DEVICE if device is None else int(device)
    if dev not in _CTX:
        _CTX[dev] = _lib.Context(dev)
    return _CTX[dev]


def multi_context(devices) -> "_lib.MultiContext":
    """One cached multi-GPU handle per device list (gsm_multi_create: contexts + a warmed NCCL communicator)."""
    key = tuple(int(d) for d in devices)
    if key not in _CTX:
        _CTX[key] = _lib.MultiContext(key)
    return _CTX[key]


# ---------------------------------------------------------------------------------------------
# fit / predict / predictprob / status
# ---------------------------------------------------------------------------------------------
def _single_column(data: GeoTable):
    names = data.names()
    return names


def _nvariables(model) -> int:
    return getattr(getattr(model, "fun", None), "nvariables", 1)


def _checkcompat(model, data: GeoTable):  # krig.jl:78-85
    nfeat = len(data.names())
    nvar = _nvariables(model)
    if isinstance(model, _KrigingModel) and nfeat != nvar:
        raise ValueError(f"{nfeat} data column(s) provided to {nvar}-variate Kriging model")


def _marginal_model(model, p):
    """the univariate model of variable p of a cokriging model with a proportional covariance"""
    f = model.fun.marginal(p)
    if isinstance(model, SimpleKriging):
        return SimpleKriging(f, float(np.atleast_1d(model.mean)[p]))
    if isinstance(model, OrdinaryKriging):
        return OrdinaryKriging(f)
    return UniversalKriging(f, exps=model.exps)


def _pointset_of(domain, centroids_ok=False) -> np.ndarray:
    """_pointsupport (models.jl:265-270): centroids of the data domain.  Data on a grid are claimed where the reference
    itself only looks at centroids: IDW / NN (idw.jl:84, nn.jl:62), and every model once fitpredict(point=true) has
    replaced the data geometries by their centroids."""
    if isinstance(domain, PointSet):
        return domain.coords
    if isinstance(domain, CartesianGrid) and centroids_ok:
        axes = [o + (np.arange(d) + 0.5) * sp for o, sp, d in zip(domain.origin, domain.spacing, domain.dims)]
        mesh = np.meshgrid(*axes, indexing="ij")
        return np.ascontiguousarray(np.stack([m_.ravel(order="F") for m_ in mesh], axis=1))
    raise NotImplementedError("GPU path claims PointSet data only")


def _encode(col):
    """(float64 codes with NaN for `missing`, categories or None): NN predicts any value type, e.g. strings
    (test/nn.jl:3-10); values cross the C ABI as the index of their category."""
    arr = col if np.ma.isMaskedArray(col) else np.asarray(col, dtype=object if isinstance(col, (list, tuple)) else None)
    if np.ma.isMaskedArray(arr) or arr.dtype.kind in "fiub":
        return np.ma.filled(np.ma.asarray(arr).astype(np.float64), np.nan), None
    miss = np.array([v is None or (isinstance(v, float) and v != v) for v in arr])
    if arr.dtype.kind == "O" and all(isinstance(v, (int, float)) and not isinstance(v, bool) for v in arr[~miss]):
        out = np.full(len(arr), np.nan)
        out[~miss] = [float(v) for v in arr[~miss]]
        return out, None
    cats, inv = np.unique(np.asarray([str(v) for v in arr[~miss]]), return_inverse=True)
    out = np.full(len(arr), np.nan)
    out[~miss] = inv
    return out, cats


def _values_of(data: "GeoTable", name) -> np.ndarray:
    """Column as float64 with `missing` as NaN (None, numpy masked entries and NaN all mean missing)."""
    col = data.table[name]
    if np.ma.isMaskedArray(col):
        return np.ma.filled(col.astype(np.float64), np.nan)
    if isinstance(col, (list, tuple)) and any(c is None for c in col):
        return np.array([np.nan if c is None else float(c) for c in col], dtype=np.float64)
    return np.asarray(col, dtype=np.float64)


def _drop_missing(model, X, z):
    """Global fit with missing values (krig.jl:182-209 + 157-161): the zeroed rows / columns leave the reduced system
    on the samples that have a value, and the untouched constraint rows of the missing samples force F_m nu = 0.
    When F_m has full column rank (always for OrdinaryKriging; for UniversalKriging as soon as enough missing samples
    are in general position) nu = 0: SimpleKriging with mean 0 on the reduced sample set.  SimpleKriging itself just
    loses the missing samples."""
    miss = np.isnan(z)
    if not miss.any():
        return model, X, z
    keep = ~miss
    if isinstance(model, SimpleKriging):
        return model, X[keep], z[keep]
    if isinstance(model, UniversalKriging):
        F = np.ones((int(miss.sum()), len(model.exps)))
        for q, e in enumerate(model.exps):
            for d_, pw in enumerate(e):
                F[:, q] *= X[miss][:, d_] ** pw
        if np.linalg.matrix_rank(F) < len(model.exps):
            raise NotImplementedError("global UniversalKriging with fewer missing samples (in general position) than "
                                      "drift functions is not claimed by the GPU path")
    return SimpleKriging(model.fun, 0.0), X[keep], z[keep]


class Dirac:
    """Distributions.Dirac: the predictive distribution of the deterministic models (idw.jl:57, nn.jl)."""

    def __init__(self, value):
        self.value = value

    def mean(self):
        return self.value

    def __repr__(self):
        return f"Dirac(value={self.value!r})"


class FittedKriging:
    def __init__(self, model, data, ctx, gfit, alpha=1.0):
        self.model, self.data, self._ctx, self._gfit, self._alpha = model, data, ctx, gfit, alpha


class FittedIDW:
    def __init__(self, model, data, ctx, alpha=1.0):
        self.model, self.data, self._ctx, self._alpha = model, data, ctx, alpha


class FittedNN:
    """fit(NN(), geotable) (nn.jl:30-41): the samples whose value is not `missing` (nn.jl:47-54 takes the nearest of
    those), values as category codes when they are not numbers."""

    def __init__(self, model, data, ctx, cats, alpha=1.0):
        self.model, self.data, self._ctx, self._cats, self._alpha = model, data, ctx, cats, alpha


class FittedCoKriging:
    """fit(model, geotable) for a cokriging model with the proportional function B * gamma0 and isotopic data: the
    reference's interleaved system (krig.jl:88-179 with nvar > 1) decouples into one univariate system per variable
    (DESIGN.md 7), fitted on demand for the column a `vars` tuple binds to that variable (krig.jl:245-258)."""

    def __init__(self, model, data, device):
        self.model, self.data, self._device = model, data, device
        self._fits = {}
        for p, name in enumerate(data.names()):
            self._fit_of(p, name)

    def _fit_of(self, p, name):
        if (p, name) not in self._fits:
            one = GeoTable({name: self.data.table[name]}, self.data.domain)
            self._fits[(p, name)] = fit(_marginal_model(self.model, p), one, device=self._device)
        return self._fits[(p, name)]


def fit(model, data: GeoTable, *, device=None, _alpha=1.0):
    """fit(model, geotable) (models.jl:38; krig.jl:58-75; idw.jl:43-49)."""
    _checkcompat(model, data)
    if _nvariables(model) > 1:
        for name in data.names():
            if np.isnan(_values_of(data, name)).any():
                raise NotImplementedError("cokriging with missing values (heterotopic data) is not claimed by the GPU path")
        return FittedCoKriging(model, data, device)
    ctx = _lib.Context(_DEFAULT_DEVICE if device is None else device)  # a fitted model owns its samples
    X = _pointset_of(data.domain, centroids_ok=isinstance(model, (IDW, NN)))
    X = X * _alpha if _alpha != 1.0 else X
    var = data.names()[0]
    if isinstance(model, NN):
        z, cats = _encode(data.table[var])
        keep = ~np.isnan(z)
        if keep.any():
            ctx.set_samples(X[keep], z[keep])
        else:
            ctx = None                                   # every value missing: predict answers `missing`
        return FittedNN(model, data, ctx, cats, _alpha)
    z = _values_of(data, var)
    if isinstance(model, IDW):
        ctx.set_samples(X, z)
        return FittedIDW(model, data, ctx, _alpha)
    model_, X, z = _drop_missing(model, X, z)      # missing values: the reduced system (see _drop_missing)
    ctx.set_samples(X, z)
    m = model_._scaled(_alpha) if _alpha != 1.0 else model_
    gfit = ctx.global_fit(m.fun._c(), m._c())
    return FittedKriging(model, data, ctx, gfit, _alpha)


def status(fitted) -> bool:
    """status(fitted) (krig.jl:49-52; idw.jl:36)."""
    if isinstance(fitted, (FittedIDW, FittedNN)):
        return True
    if isinstance(fitted, FittedCoKriging):
        return all(f._gfit.status() for f in fitted._fits.values())
    return fitted._gfit.status()


def _as_point(g, dim):
    p = np.asarray(g.coords if isinstance(g, PointSet) else g, dtype=np.float64).reshape(1, -1)
    if p.shape[1] != dim:
        raise ValueError("point has the wrong dimension")
    return p


def _check_vars(fitted, var):
    """Returns (name, scalar?) and mirrors the ArgumentError of models.jl:56-57 / 71-72."""
    if isinstance(var, str):
        names = (var,)
        scalar = True
    else:
        names = tuple(var)
        scalar = False
        if len(names) > 1:
            raise ValueError("cannot use univariate model to predict multiple variables")
    if names[0] not in fitted.data.names():
        raise KeyError(names[0])
    return names[0], scalar


def _target_of(fitted, g):
    """point target at the centroid of g, and the block offsets when g has a volume (krig.jl:316-361 on a geometry)"""
    block = None
    if isinstance(g, Box):
        if isinstance(fitted, FittedKriging):
            block = block_points(g, fitted.model._range()) * fitted._alpha
        g = g.centroid()
    p = _as_point(g, fitted._ctx.dim) * fitted._alpha
    return _lib.Context.point_targets(p), block


def _krige_at(fitted, t, block, want_var):
    fitted._ctx.set_block_support(block)
    try:
        return fitted._gfit.predict(t, want_var=want_var)
    finally:
        if block is not None:
            fitted._ctx.set_block_support(None)


def _nn_at(fitted, g):
    """value of the nearest sample with a value (nn.jl:47-54); None stands for `missing`"""
    if fitted._ctx is None:
        return None
    g = g.centroid() if isinstance(g, Box) else g
    p = _as_point(g, fitted._ctx.dim) * fitted._alpha
    code = float(fitted._ctx.nn(_lib.Context.point_targets(p))[0])
    return code if fitted._cats is None else str(fitted._cats[int(code)])


def _co_vars(fitted, var):
    """the columns bound to the variables of a cokriging model: krigmean indexes vars[p] for p in 1:nvar (krig.jl:250-256)"""
    names = (var,) if isinstance(var, str) else tuple(var)
    if len(names) != _nvariables(fitted.model):
        raise ValueError(f"{len(names)} variable(s) requested from a {_nvariables(fitted.model)}-variate Kriging model")
    for nm in names:
        if nm not in fitted.data.names():
            raise KeyError(nm)
    return names


def _co_predict(fitted, var, g, want_var):
    names = _co_vars(fitted, var)
    mu, va = [], []
    for p, nm in enumerate(names):
        f = fitted._fit_of(p, nm)
        t, block = _target_of(f, g)
        m_, v_ = _krige_at(f, t, block, want_var)
        mu.append(float(m_[0]))
        va.append(float(v_[0]) if want_var else 0.0)
    return np.array(mu), np.array(va)


def predict(fitted, var, g):
    """predict(fitted, var, g) (krig.jl:215-219; idw.jl:55); g: a point or a Box (change of support)."""
    if isinstance(fitted, FittedCoKriging):
        return list(_co_predict(fitted, var, g, False)[0])
    if isinstance(fitted, FittedNN):
        _, scalar = _check_vars(fitted, var)
        v = _nn_at(fitted, g)
        return v if scalar else [v]
    _, scalar = _check_vars(fitted, var)
    t, block = _target_of(fitted, g)
    if isinstance(fitted, FittedIDW):
        mean, _ = fitted._ctx.idw(fitted.model.exponent, t)
        if mean[0] != mean[0]:                          # a `missing` value among the samples (idw.jl:59-73)
            return None if scalar else [None]
    else:
        mean, _ = _krige_at(fitted, t, block, False)
    return float(mean[0]) if scalar else [float(mean[0])]


def predictprob(fitted, var, g):
    """predictprob(fitted, var, g): Normal(mu, sqrt(var)) for Kriging (krig.jl:223-229)."""
    if isinstance(fitted, FittedNN):
        _, scalar = _check_vars(fitted, var)
        d = Dirac(_nn_at(fitted, g))                     # predictprob = Dirac(predict) (nn.jl:45)
        return d if scalar else [d]
    if isinstance(fitted, FittedCoKriging):
        # Sigma = B sigma0^2 + 1e-10 I (krig.jl:264-293 for B * gamma0): every entry scales the base variance; the
        # diagonal is each variable's own solve, the off-diagonal entries take sigma0^2 from the first variable
        mu, va = _co_predict(fitted, var, g, True)
        B = fitted.model.fun.B
        s0 = (va[0] - 1e-10) / B[0, 0]
        Sigma = B * s0 + 1e-10 * np.eye(len(mu))
        Sigma[np.diag_indices(len(mu))] = va
        return MvNormal(mu, Sigma)
    _, scalar = _check_vars(fitted, var)
    t, block = _target_of(fitted, g)
    if isinstance(fitted, FittedIDW):
        mean, _ = fitted._ctx.idw(fitted.model.exponent, t)      # predictprob = Dirac(predict) (idw.jl:57)
        d = Dirac(float(mean[0]))
        return d if scalar else [d]
    mean, var_ = _krige_at(fitted, t, block, True)
    if not scalar:                                        # a tuple of variables: MvNormal(mu, Sigma), krig.jl:231-237
        return MvNormal([float(mean[0])], [[float(var_[0])]])
    if var_[0] < 0:
        raise ValueError("DomainError: sqrt of negative variance")  # krig.jl:228
    return Normal(float(mean[0]), math.sqrt(float(var_[0])))


# ---------------------------------------------------------------------------------------------
# fitpredict
# ---------------------------------------------------------------------------------------------
def _scale_factor(model, data_dom, dom) -> float:
    """alpha of `_scale` (models.jl:272-284)."""
    r = model._range()
    a1 = 1.0 if math.isinf(r) else r
    return 1.0 / max(a1, data_dom._absmax(), dom._absmax())


def _targets_for(dom, alpha, first=0, count=-1):
    if isinstance(dom, CartesianGrid):
        origin = tuple(alpha * np.float64(o) for o in dom.origin)
        spacing = tuple(alpha * np.float64(s) for s in dom.spacing)
        return _lib.Context.grid_targets(origin, spacing, dom.dims, first, count)
    return _lib.Context.point_targets(dom.coords * alpha, first=first, count=count)


def block_points(dom, vrange: float) -> np.ndarray:
    """Offsets (nq, dim) from the centroid of the points that stand for one cell of the grid ``dom`` (or for a ``Box``)
    under change of support.

    Which points represent a geometry is decided upstream in GeoStatsFunctions (``f(point, geometry)`` behind
    ``pairwise!``, krig.jl:125,350), not in the reference tree.  The rule restated here is the recalled one: spacing
    ``s = min(range, shortest side) / 3``, ``n_d = ceil(side_d / s)`` points per dimension on the regular lattice that
    includes the cell boundary (Meshes ``RegularSampling`` on a non-periodic parametrisation).  The only thing the
    reference's tests pin about block support is ``var >= 0`` (test/krig.jl:221-242), so parity of the rule itself is
    unpinned; a caller who knows better passes ``block_offsets=`` to fitpredict."""
    sides = np.abs(np.asarray(dom.sides if isinstance(dom, Box) else dom.spacing, dtype=np.float64))
    pos = sides[sides > 0]
    short = float(pos.min()) if len(pos) else 1.0
    s = (min(float(vrange), short) if (vrange > 0 and math.isfinite(vrange)) else short) / 3.0
    axes = []
    for side in sides:
        n = max(int(math.ceil(side / s - 1e-9)), 1) if side > 0 else 1
        axes.append(np.linspace(-0.5 * side, 0.5 * side, n) if n > 1 else np.zeros(1))
    mesh = np.meshgrid(*axes, indexing="ij")
    # first coordinate fastest, like the element order of the grid itself
    return np.stack([m_.ravel(order="F") for m_ in mesh], axis=1)


def fitpredict(model, data: GeoTable, dom, *, point=True, prob=False, neighbors=True, minneighbors=1,
               maxneighbors=10, neighborhood=None, distance="euclidean", device=None, devices=None, first=0,
               count=-1, ctx=None, out=None, block_offsets=None, _device_samples=None):
    """fitpredict(model, geotable, domain; options) (models.jl:102-129).

    ``first``/``count`` (not in the reference) select a contiguous shard of the target index for
    multi-GPU runs: every rank predicts its slab of the same domain (models.jl:236-247 chunking).
    ``out`` may hold preallocated (pinned) arrays ``mean`` / ``var`` / ``status`` the results DMA into.
    ``devices=[0, 1, ...]`` (not in the reference) runs the call on several GPUs of the box through the multi-GPU
    C ABI (gsm_multi_*): samples replicated by one NCCL broadcast, target slabs, one host thread per device."""
    if distance != "euclidean":
        raise NotImplementedError("only the Euclidean distance is claimed by the GPU path")
    if not isinstance(dom, (PointSet, CartesianGrid)):
        dom = PointSet(dom)
    # change of support (point=false, models.jl:114-115,136): the elements of a PointSet are points anyway, and IDW / NN
    # only look at centroids (idw.jl:82, nn.jl); a Kriging model on a grid averages the right-hand side over the cell
    block = None
    if not point and isinstance(dom, CartesianGrid) and isinstance(model, _KrigingModel):
        block = block_points(dom, model._range()) if block_offsets is None else np.asarray(block_offsets, dtype=np.float64)
        if block.ndim != 2 or block.shape[1] != dom.dim or len(block) < 1:
            raise ValueError("block_offsets must be (nq, dim) with nq >= 1")
        if len(block) > _lib.GSM_MAX_BLOCK_POINTS:
            raise NotImplementedError("change of support with more than GSM_MAX_BLOCK_POINTS points per cell is not claimed")
    if _nvariables(model) > 1:
        # CoKriging with a proportional covariance B * gamma0 and isotopic data: one univariate solve per variable
        _checkcompat(model, data)
        cols = {}
        for p_, name in enumerate(data.names()):
            if np.isnan(_values_of(data, name)).any():
                raise NotImplementedError("cokriging with missing values (heterotopic data) is not claimed by the GPU path")
            one = fitpredict(_marginal_model(model, p_), GeoTable({name: data.table[name]}, data.domain), dom, point=point,
                             prob=prob, neighbors=neighbors, minneighbors=minneighbors, maxneighbors=maxneighbors,
                             neighborhood=neighborhood, distance=distance, device=device, devices=devices, first=first,
                             count=count, ctx=ctx, block_offsets=block_offsets)
            cols[name] = one.table[name]
        pred = GeoTable.__new__(GeoTable)
        pred.table = cols
        pred.domain = dom
        pred.shard = one.shard
        return pred
    if _device_samples is not None:
        # (distributed.fitpredict_sharded) the samples already sit in device memory on this rank: an NCCL broadcast
        # delivered them; `data` only carries the column name
        coords_d, values_d, nobs, sdim, absmax_hint, names = _device_samples
    else:
        _checkcompat(model, data)
        names = data.names()
        if len(names) != 1:
            raise NotImplementedError("GPU path claims univariate tables only")
        cats = None
        if isinstance(model, NN):
            z, cats = _encode(data.table[names[0]])
        else:
            z = _values_of(data, names[0])
        # _pointsupport (models.jl:115): centroids of the data geometries; with point=false the data keep their
        # geometries, which only IDW / NN reduce to centroids themselves
        X = _pointset_of(data.domain, centroids_ok=point or isinstance(model, (IDW, NN)))
        if isinstance(model, NN) and not neighbors and np.isnan(z).any():
            # fitpredictfull: the nearest sample WITH a value (nn.jl:47-54) = nearest of the samples that have one
            # (the scale factor still comes from all samples, models.jl:286-291)
            keep = ~np.isnan(z)
            if not keep.any():
                raise NotImplementedError("NN on a table without values is not claimed by the GPU path")
            _nn_floor = float(np.abs(X).max())
            X, z = np.ascontiguousarray(X[keep]), z[keep]
        nobs = len(z)
    if ctx is None and devices is not None and len(devices) > 1:
        ctx = multi_context(devices)
    ctx = ctx or default_context(device if devices is None else devices[0])
    multi = isinstance(ctx, _lib.MultiContext)
    # _scale (models.jl:272-291) on the device: the raw coordinates are uploaded as they are (straight DMA when the
    # caller's arrays are page-locked), absmax / alpha / `dat |> Scale(alpha)` / the NaN check run as kernels
    r = model._range()
    floor = max(1.0 if math.isinf(r) else r, dom._absmax())
    if _device_samples is None and isinstance(model, NN) and not neighbors and len(z) != data.domain.nelements():
        floor = max(floor, _nn_floor)
    kmodel = model
    if _device_samples is not None:
        alpha, nnan = ctx.set_samples_scaled(coords_d, values_d, floor, dim=sdim, n=nobs)
    else:
        if not neighbors and isinstance(model, _KrigingModel) and np.isnan(z).any():
            # fitpredictfull with missing values: the reduced system on the host side (alpha still from ALL samples)
            floor = max(floor, float(np.abs(X).max()))
            kmodel, X, z = _drop_missing(model, X, z)
        alpha, nnan = ctx.set_samples_scaled(X, z, floor)
    if nnan and not ((isinstance(model, (_KrigingModel, NN)) and neighbors) or isinstance(model, IDW)):
        raise NotImplementedError("missing values are not claimed for this model / option combination")
    out = out or {}
    t = _targets_for(dom, alpha, first, count)
    radius = 0.0
    if neighborhood is not None:
        radius = alpha * float(neighborhood)           # sneigh = alpha * neigh (models.jl:281), ball radius
    M = _lib.Context.target_count(t)
    if isinstance(model, NN):
        if multi:
            raise NotImplementedError("NN is served by the single-device context")
        # nearest of the k nearest = nearest of all (nn.jl:47-54) -- unless a ball neighbourhood leaves fewer than
        # minneighbors samples inside the ball: that target is `missing` (models.jl:167-181)
        if neighbors and nnan:
            # a column with missings behind a neighbour search: the nearest of the k nearest that HAS a value, `missing`
            # when none has (models.jl:162-184 with nn.jl:47-54)
            col, st = ctx.nn_local(t, _clamp_max(maxneighbors, nobs), minneighbors, radius, out=out.get("mean"))
            col = _mask_missing(col, st)
        else:
            col = ctx.nn(t, out=out.get("mean"))
        if neighbors and neighborhood is not None and not nnan:
            _, cnt = ctx.knn(t, _clamp_max(maxneighbors, nobs), radius)
            kmax = _clamp_max(maxneighbors, nobs)
            minn = 1 if (minneighbors > kmax or minneighbors < 1) else minneighbors
            if (cnt < minn).any():
                col = np.ma.masked_array(col, mask=cnt < minn)
        if _device_samples is None and cats is not None:    # category codes back to the values
            codes = np.ma.filled(col, 0).astype(np.int64)
            vals = np.asarray(cats, dtype=object)[codes]
            col = np.ma.masked_array(vals, mask=np.ma.getmaskarray(col)) if np.ma.isMaskedArray(col) else vals
        if prob:
            col = Dirac(col)                               # predictprob = Dirac(predict) (nn.jl)
    elif isinstance(model, IDW):
        if neighbors:
            mean, st = ctx.idw(model.exponent, t, maxneighbors=_clamp_max(maxneighbors, nobs),
                               minneighbors=minneighbors, radius=radius)
        else:
            mean, st = ctx.idw(model.exponent, t)
        col = _mask_missing(mean, st)
        if nnan:
            # a `missing` value among the weighted samples makes the sum `missing` (idw.jl:59-73): NaN on the device
            nanmask = np.isnan(np.ma.getdata(col))
            if nanmask.any():
                col = np.ma.masked_array(np.ma.getdata(col), mask=nanmask | np.ma.getmaskarray(col))
        if prob:
            col = Dirac(col)                               # predictprob = Dirac(predict) (idw.jl:57)
    else:
        sm = kmodel._scaled(alpha)
        ctx.set_block_support(None if block is None else block * alpha)
        if neighbors:
            mean, var, st, _ = ctx.local_krige(sm.fun._c(), sm._c(), t, _clamp_max(maxneighbors, nobs),
                                               minneighbors, radius, want_var=prob, out_mean=out.get("mean"),
                                               out_var=out.get("var"), out_status=out.get("status"))
        elif multi:
            mean, var, fit_ok = ctx.global_krige(sm.fun._c(), sm._c(), t, want_var=prob)
            st = np.zeros(M, dtype=np.uint8)
            if not fit_ok:
                st[:] = _lib.GSM_ST_SINGULAR           # status(fitted) == false (krig.jl:49-52)
        else:
            gfit = ctx.global_fit(sm.fun._c(), sm._c())
            mean, var = gfit.predict(t, want_var=prob)
            st = np.zeros(M, dtype=np.uint8)
            if not gfit.status():
                st[:] = _lib.GSM_ST_SINGULAR           # status(fitted) == false (krig.jl:49-52)
            gfit.close()
        if block is not None:
            ctx.set_block_support(None)
        # the library reduces the status / variance columns on the device (gsm_get_summary): the host only
        # scans them (tens of MB) when that is not available
        sm = ctx.summary() if neighbors else {"total": -1}
        known = sm["total"] == M
        clean = (sm["missing"] == 0 and sm["singular"] == 0) if known else not st.any()
        if prob:
            if known:
                bad = sm["nonpositive_var"] > 0
            else:
                bad = bool(np.any(~(var > 0) & (st == 0)))     # (NaN counts as not positive)
            if bad:
                raise ValueError("PosDefException: non-positive Kriging variance")  # MvNormal(mu, Sigma), krig.jl:236
            # Normal.(mean, var), models.jl:298-299
            col = Normal(mean, var) if clean else Normal(_mask_missing(mean, st), _mask_missing(var, st))
        else:
            col = mean if clean else _mask_missing(mean, st)
    out_dom = dom if (first == 0 and count < 0) else None
    pred = GeoTable.__new__(GeoTable)
    pred.table = {names[0]: col}
    pred.domain = out_dom if out_dom is not None else dom   # georef over the ORIGINAL domain, models.jl:128
    pred.shard = (first, M)
    return pred


def _clamp_max(maxneighbors, nobs):
    return nobs if (maxneighbors > nobs or maxneighbors < 1) else maxneighbors   # models.jl:144-147


def _mask_missing(a, st):
    miss = st == _lib.GSM_ST_MISSING
    if miss.any():
        return np.ma.masked_array(a, mask=miss)   # `missing` predictions, models.jl:178-181
    return a
