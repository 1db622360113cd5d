"""ctypes binding of libgsm_b200.so -- the same C ABI a Julia ``ccall`` binds (include/gsm_b200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is present every
compute entry point raises.  Only the structure definitions and symbol table are usable without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

GSM_MAX_DIM = 3
GSM_MAX_DRIFT = 10
GSM_MAX_NEIGHBORS = 64

GSM_OK, GSM_ERR_INVALID, GSM_ERR_CUDA, GSM_ERR_NO_SAMPLES, GSM_ERR_UNSUPPORTED, GSM_ERR_SINGULAR = range(6)
GSM_ST_OK, GSM_ST_MISSING, GSM_ST_SINGULAR = 0, 1, 2
GSM_VARIO_GAUSSIAN, GSM_VARIO_SPHERICAL, GSM_VARIO_EXPONENTIAL = 0, 1, 2
GSM_VARIO_CUBIC, GSM_VARIO_PENTASPHERICAL, GSM_VARIO_SINEHOLE, GSM_VARIO_CIRCULAR, GSM_VARIO_NUGGET = 3, 4, 5, 6, 7
GSM_MAX_STRUCTURES = 4
GSM_MAX_BLOCK_POINTS = 512
GSM_KRIG_SIMPLE, GSM_KRIG_ORDINARY, GSM_KRIG_UNIVERSAL = 0, 1, 2
GSM_TARGETS_GRID, GSM_TARGETS_POINTS = 0, 1

EXPORTED_SYMBOLS = [
    "gsm_abi_version", "gsm_create", "gsm_destroy", "gsm_last_error", "gsm_get_timings", "gsm_get_summary", "gsm_synchronize",
    "gsm_host_alloc", "gsm_host_free", "gsm_set_samples", "gsm_set_samples_scaled", "gsm_knn", "gsm_local_krige",
    "gsm_global_krige_fit", "gsm_global_krige_status", "gsm_global_krige_predict", "gsm_global_krige_free",
    "gsm_idw", "gsm_nn", "gsm_nn_local", "gsm_measure_fp64_peak", "gsm_set_option", "gsm_set_block_support",
    "gsm_multi_create", "gsm_multi_destroy", "gsm_multi_last_error", "gsm_multi_device_count", "gsm_multi_context",
    "gsm_multi_get_timings", "gsm_multi_get_summary", "gsm_multi_set_samples", "gsm_multi_set_samples_scaled",
    "gsm_multi_slab", "gsm_multi_local_krige", "gsm_multi_idw", "gsm_multi_global_krige",
]


class Structure(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("sill", C.c_double), ("nugget", C.c_double),
                ("range", C.c_double), ("metric", C.c_double * 9)]


class Variogram(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nstruct", C.c_int32), ("sill", C.c_double), ("nugget", C.c_double),
                ("range", C.c_double), ("structs", Structure * 4)]


class Model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ndrift", C.c_int32), ("mean", C.c_double),
                ("exps", (C.c_int32 * GSM_MAX_DIM) * GSM_MAX_DRIFT)]


class Targets(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("origin", C.c_double * GSM_MAX_DIM),
                ("spacing", C.c_double * GSM_MAX_DIM), ("dims", C.c_int64 * GSM_MAX_DIM),
                ("points", C.c_void_p), ("npoints", C.c_int64), ("first", C.c_int64), ("count", C.c_int64)]


class NeighborOpts(C.Structure):
    _fields_ = [("maxneighbors", C.c_int32), ("minneighbors", C.c_int32), ("radius", C.c_double)]


class Summary(C.Structure):
    _fields_ = [("missing", C.c_int64), ("singular", C.c_int64), ("nonpositive_var", C.c_int64), ("total", C.c_int64)]


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("bin_ms", C.c_double), ("knn_ms", C.c_double),
                ("solve_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double),
                ("launches", C.c_int64), ("index_ms", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class MultiTimings(C.Structure):
    _fields_ = [("upload_ms", C.c_double), ("bcast_ms", C.c_double), ("bins_ms", C.c_double),
                ("predict_ms", C.c_double), ("wall_ms", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class GsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libgsm_b200 error {code}: {msg}")
        self.code = code


_LIB = None


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libgsm_b200.so")


def load():
    """Load libgsm_b200.so (fails loudly when it has not been built)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    lib.gsm_abi_version.restype = C.c_int
    lib.gsm_create.argtypes = [C.c_int, P(vp)]
    lib.gsm_destroy.argtypes = [vp]
    lib.gsm_last_error.argtypes = [vp]
    lib.gsm_last_error.restype = C.c_char_p
    lib.gsm_get_timings.argtypes = [vp, P(Timings)]
    lib.gsm_get_summary.argtypes = [vp, P(Summary)]
    lib.gsm_synchronize.argtypes = [vp]
    lib.gsm_host_alloc.argtypes = [P(vp), i64]
    lib.gsm_host_free.argtypes = [vp]
    lib.gsm_set_samples.argtypes = [vp, vp, i32, i64, vp]
    lib.gsm_set_samples_scaled.argtypes = [vp, vp, i32, i64, vp, C.c_double, P(C.c_double), P(C.c_int64)]
    lib.gsm_knn.argtypes = [vp, P(Targets), P(NeighborOpts), vp, vp]
    lib.gsm_local_krige.argtypes = [vp, P(Variogram), P(Model), P(Targets), P(NeighborOpts), vp, vp, vp, vp]
    lib.gsm_global_krige_fit.argtypes = [vp, P(Variogram), P(Model), P(vp)]
    lib.gsm_global_krige_status.argtypes = [vp]
    lib.gsm_global_krige_predict.argtypes = [vp, P(Targets), vp, vp]
    lib.gsm_global_krige_free.argtypes = [vp]
    lib.gsm_idw.argtypes = [vp, dbl, P(Targets), P(NeighborOpts), vp, vp]
    lib.gsm_nn.argtypes = [vp, P(Targets), vp]
    lib.gsm_nn_local.argtypes = [vp, P(Targets), P(NeighborOpts), vp, vp]
    lib.gsm_measure_fp64_peak.argtypes = [vp, P(dbl), P(dbl)]
    lib.gsm_set_option.argtypes = [vp, C.c_char_p, C.c_char_p]
    lib.gsm_set_block_support.argtypes = [vp, C.c_void_p, C.c_int32, C.c_int32]
    lib.gsm_multi_create.argtypes = [P(i32), i32, P(vp)]
    lib.gsm_multi_destroy.argtypes = [vp]
    lib.gsm_multi_last_error.argtypes = [vp]
    lib.gsm_multi_last_error.restype = C.c_char_p
    lib.gsm_multi_device_count.argtypes = [vp]
    lib.gsm_multi_context.argtypes = [vp, i32]
    lib.gsm_multi_context.restype = vp
    lib.gsm_multi_get_timings.argtypes = [vp, P(MultiTimings)]
    lib.gsm_multi_get_summary.argtypes = [vp, P(Summary)]
    lib.gsm_multi_set_samples.argtypes = [vp, vp, i32, i64, vp]
    lib.gsm_multi_set_samples_scaled.argtypes = [vp, vp, i32, i64, vp, C.c_double, P(C.c_double), P(C.c_int64)]
    lib.gsm_multi_slab.argtypes = [vp, P(Targets), i32, P(i64), P(i64)]
    lib.gsm_multi_local_krige.argtypes = [vp, P(Variogram), P(Model), P(Targets), P(NeighborOpts), vp, vp, vp]
    lib.gsm_multi_idw.argtypes = [vp, dbl, P(Targets), P(NeighborOpts), vp, vp]
    lib.gsm_multi_global_krige.argtypes = [vp, P(Variogram), P(Model), P(Targets), vp, vp, P(i32)]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("gsm_last_error", "gsm_multi_last_error", "gsm_multi_context"):
            fn.restype = C.c_int
    _LIB = lib
    return lib


def _ptr(a):
    """Raw address of a numpy array, an int (device pointer), a torch tensor, or None."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")


class _PinnedPool:
    """Result arrays in page-locked memory without asking the caller for them: buffers are recycled when the last
    numpy view of a result dies (finalizer on the ctypes buffer every view keeps alive), so steady-state calls DMA
    straight into pinned memory and pay no cudaHostAlloc."""

    CAP = 2 << 30   # pooled (idle) bytes kept at most

    def __init__(self):
        self._free, self._idle = {}, 0

    def empty(self, shape, dtype):
        import weakref
        dtype = np.dtype(dtype)
        shape = tuple(int(x) for x in np.atleast_1d(shape))
        count = int(np.prod(shape))
        key = max(4096, -(-count * dtype.itemsize // 4096) * 4096)
        lst = self._free.get(key)
        if lst:
            ptr = lst.pop()
            self._idle -= key
        else:
            p = C.c_void_p()
            rc = load().gsm_host_alloc(C.byref(p), key)
            if rc != GSM_OK:
                return np.empty(shape, dtype)   # pinning failed (limits): pageable memory still works
            ptr = p.value
        buf = (C.c_char * key).from_address(ptr)
        weakref.finalize(buf, self._release, key, ptr)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def _release(self, key, ptr):
        try:
            if self._idle + key <= self.CAP:
                self._free.setdefault(key, []).append(ptr)
                self._idle += key
            else:
                load().gsm_host_free(C.c_void_p(ptr))
        except Exception:
            pass


_pool = _PinnedPool()


class PinnedArray:
    """A numpy view over page-locked host memory from gsm_host_alloc (outputs DMA straight into it)."""

    def __init__(self, shape, dtype):
        lib = load()
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        rc = lib.gsm_host_alloc(C.byref(p), nbytes)
        if rc != GSM_OK:
            raise GsmError(rc, (lib.gsm_last_error(None) or b"").decode())
        self._p = p
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p is not None:
            self.array = None
            load().gsm_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU, one stream, one sample set: the handle a Julia finalizer would own."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = C.c_void_p()
        rc = self._lib.gsm_create(int(device), C.byref(h))
        if rc != GSM_OK:
            raise GsmError(rc, (self._lib.gsm_last_error(None) or b"").decode())
        self._h = h
        self.device = int(device)
        self.dim = 0
        self.n = 0

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lib.gsm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != GSM_OK:
            raise GsmError(rc, (self._lib.gsm_last_error(self._h) or b"").decode())

    def timings(self) -> dict:
        t = Timings()
        self._check(self._lib.gsm_get_timings(self._h, C.byref(t)))
        return t.asdict()

    def summary(self) -> dict:
        """status counts of the last neighbourhood call, reduced on the device (total = -1: not available)"""
        sm = Summary()
        self._check(self._lib.gsm_get_summary(self._h, C.byref(sm)))
        return {k: getattr(sm, k) for k, _ in sm._fields_}

    def synchronize(self):
        self._check(self._lib.gsm_synchronize(self._h))

    def set_option(self, key: str, value) -> None:
        """developer options (gsm_set_option): local_impl, shr_smax, bin_occupancy"""
        self._check(self._lib.gsm_set_option(self._h, str(key).encode(), str(value).encode()))

    def set_block_support(self, offsets=None) -> None:
        """gsm_set_block_support: the (nq, dim) points, relative to the centroid and in sample coordinates, that stand
        for every target of the following Kriging calls; None or an empty table returns to point support."""
        if offsets is None or len(offsets) == 0:
            self._check(self._lib.gsm_set_block_support(self._h, None, 0, 0))
            return
        off = np.ascontiguousarray(offsets, dtype=np.float64)
        if off.ndim == 1:
            off = off[:, None]
        self._check(self._lib.gsm_set_block_support(self._h, off.ctypes.data, off.shape[1], off.shape[0]))

    def set_samples(self, coords, values, dim=None, n=None):
        """coords: (n, dim) float64 C-order numpy array, or a device pointer / torch tensor with dim, n."""
        if isinstance(coords, np.ndarray):
            coords = np.ascontiguousarray(coords, dtype=np.float64)
            if coords.ndim == 1:
                coords = coords[:, None]
            n, dim = coords.shape
            values = np.ascontiguousarray(values, dtype=np.float64)
            if values.shape != (n,):
                raise ValueError("values must have one entry per sample")
        self._keep = (coords, values)
        self._check(self._lib.gsm_set_samples(self._h, _ptr(coords), int(dim), int(n), _ptr(values)))
        self.dim, self.n = int(dim), int(n)

    def set_samples_scaled(self, coords, values, scale_floor, dim=None, n=None):
        """Upload RAW coordinates; absmax, alpha = 1 / max(scale_floor, absmax) and the scaling happen on the device
        (gsm_set_samples_scaled).  Returns (alpha, number of NaN values).  coords / values: numpy arrays, or device
        pointers / torch CUDA tensors with dim and n (samples that arrived by a broadcast)."""
        if isinstance(coords, np.ndarray):
            coords = np.ascontiguousarray(coords, dtype=np.float64)
            if coords.ndim == 1:
                coords = coords[:, None]
            n, dim = coords.shape
            values = np.ascontiguousarray(values, dtype=np.float64)
            if values.shape != (n,):
                raise ValueError("values must have one entry per sample")
        self._keep = (coords, values)
        alpha, nan = C.c_double(0.0), C.c_int64(0)
        self._check(self._lib.gsm_set_samples_scaled(self._h, _ptr(coords), int(dim), int(n), _ptr(values),
                                                     C.c_double(float(scale_floor)), C.byref(alpha), C.byref(nan)))
        self.dim, self.n = int(dim), int(n)
        return float(alpha.value), int(nan.value)

    # ---- descriptors -----------------------------------------------------------------------
    @staticmethod
    def grid_targets(origin, spacing, dims, first=0, count=-1) -> Targets:
        t = Targets()
        t.kind = GSM_TARGETS_GRID
        t.dim = len(dims)
        for d in range(len(dims)):
            t.origin[d] = float(origin[d])
            t.spacing[d] = float(spacing[d])
            t.dims[d] = int(dims[d])
        t.points = None
        t.npoints = 0
        t.first, t.count = int(first), int(count)
        return t

    @staticmethod
    def point_targets(points, dim=None, npoints=None, first=0, count=-1) -> Targets:
        t = Targets()
        t.kind = GSM_TARGETS_POINTS
        if isinstance(points, np.ndarray):
            points = np.ascontiguousarray(points, dtype=np.float64)
            if points.ndim == 1:
                points = points[:, None]
            npoints, dim = points.shape
        t._keep = points
        t.dim = int(dim)
        t.points = _ptr(points)
        t.npoints = int(npoints)
        t.first, t.count = int(first), int(count)
        return t

    @staticmethod
    def target_count(t: Targets) -> int:
        total = t.npoints if t.kind == GSM_TARGETS_POINTS else int(np.prod([t.dims[d] for d in range(t.dim)]))
        return total - t.first if t.count < 0 else t.count

    # ---- entry points ------------------------------------------------------------------------
    def knn(self, targets: Targets, maxneighbors: int, radius: float = 0.0):
        M = self.target_count(targets)
        k = min(maxneighbors, self.n) if 1 <= maxneighbors <= self.n else self.n
        idx = np.empty((M, k), dtype=np.int32)
        cnt = np.empty(M, dtype=np.int32)
        o = NeighborOpts(int(maxneighbors), 1, float(radius))
        self._check(self._lib.gsm_knn(self._h, C.byref(targets), C.byref(o), _ptr(idx), _ptr(cnt)))
        return idx, cnt

    def local_krige(self, vario: Variogram, model: Model, targets: Targets, maxneighbors: int, minneighbors: int = 1,
                    radius: float = 0.0, want_var=True, want_neighbors=False, out_mean=None, out_var=None,
                    out_status=None):
        M = self.target_count(targets)
        mean = _pool.empty(M, np.float64) if out_mean is None else out_mean
        var = (_pool.empty(M, np.float64) if out_var is None else out_var) if want_var else None
        status = _pool.empty(M, np.uint8) if out_status is None else out_status
        nbr = None
        if want_neighbors:
            k = min(maxneighbors, self.n) if 1 <= maxneighbors <= self.n else self.n
            nbr = np.empty((M, k), dtype=np.int32)
        o = NeighborOpts(int(maxneighbors), int(minneighbors), float(radius))
        self._check(self._lib.gsm_local_krige(self._h, C.byref(vario), C.byref(model), C.byref(targets), C.byref(o),
                                              _ptr(mean), _ptr(var), _ptr(status), _ptr(nbr)))
        return mean, var, status, nbr

    def idw(self, exponent: float, targets: Targets, maxneighbors: int = 0, minneighbors: int = 1,
            radius: float = 0.0, out_mean=None):
        M = self.target_count(targets)
        mean = _pool.empty(M, np.float64) if out_mean is None else out_mean
        status = _pool.empty(M, np.uint8)
        o = NeighborOpts(int(maxneighbors), int(minneighbors), float(radius)) if maxneighbors != 0 else None
        self._check(self._lib.gsm_idw(self._h, float(exponent), C.byref(targets),
                                      C.byref(o) if o is not None else None, _ptr(mean), _ptr(status)))
        return mean, status

    def nn(self, targets: Targets, out=None):
        """value of the nearest sample for every target (gsm_nn)"""
        M = self.target_count(targets)
        val = _pool.empty(M, np.float64) if out is None else out
        self._check(self._lib.gsm_nn(self._h, C.byref(targets), _ptr(val)))
        return val

    def nn_local(self, targets: Targets, maxneighbors: int, minneighbors: int = 1, radius: float = 0.0, out=None):
        """nearest sample WITH a value among the maxneighbors nearest (gsm_nn_local); status 1 = `missing`"""
        M = self.target_count(targets)
        val = _pool.empty(M, np.float64) if out is None else out
        status = _pool.empty(M, np.uint8)
        o = NeighborOpts(int(maxneighbors), int(minneighbors), float(radius))
        self._check(self._lib.gsm_nn_local(self._h, C.byref(targets), C.byref(o), _ptr(val), _ptr(status)))
        return val, status

    def global_fit(self, vario: Variogram, model: Model):
        h = C.c_void_p()
        self._check(self._lib.gsm_global_krige_fit(self._h, C.byref(vario), C.byref(model), C.byref(h)))
        return GlobalFit(self, h)

    def measure_fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(self._lib.gsm_measure_fp64_peak(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


class MultiContext:
    """All GPUs of a box behind ONE handle (gsm_multi_*): samples uploaded once and replicated by an NCCL broadcast,
    the target range split into contiguous slabs, one host thread per device, outputs written straight into the
    caller's host arrays.  Mirrors the Context methods fitpredict needs."""

    def __init__(self, devices):
        self._lib = load()
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self._lib.gsm_multi_create(devs, len(devices), C.byref(h))
        if rc != GSM_OK:
            raise GsmError(rc, (self._lib.gsm_multi_last_error(None) or b"").decode())
        self._h = h
        self.devices = [int(d) for d in devices]
        self.dim = 0
        self.n = 0

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lib.gsm_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != GSM_OK:
            raise GsmError(rc, (self._lib.gsm_multi_last_error(self._h) or b"").decode())

    def timings(self) -> dict:
        t = MultiTimings()
        self._check(self._lib.gsm_multi_get_timings(self._h, C.byref(t)))
        return t.asdict()

    def device_timings(self, i: int) -> dict:
        t = Timings()
        rc = self._lib.gsm_get_timings(self._lib.gsm_multi_context(self._h, int(i)), C.byref(t))
        self._check(rc)
        return t.asdict()

    def summary(self) -> dict:
        sm = Summary()
        self._check(self._lib.gsm_multi_get_summary(self._h, C.byref(sm)))
        return {k: getattr(sm, k) for k, _ in sm._fields_}

    def set_option(self, key, value) -> None:
        for i in range(len(self.devices)):
            rc = self._lib.gsm_set_option(self._lib.gsm_multi_context(self._h, i), str(key).encode(), str(value).encode())
            self._check(rc)

    def set_block_support(self, offsets=None) -> None:
        nq = 0 if offsets is None else len(offsets)
        off = np.ascontiguousarray(offsets, dtype=np.float64).reshape(nq, -1) if nq else None
        for i in range(len(self.devices)):
            rc = self._lib.gsm_set_block_support(self._lib.gsm_multi_context(self._h, i),
                                                 off.ctypes.data if nq else None, off.shape[1] if nq else 0, nq)
            self._check(rc)

    def slab(self, targets: Targets, i: int):
        f, c = C.c_int64(), C.c_int64()
        self._check(self._lib.gsm_multi_slab(self._h, C.byref(targets), int(i), C.byref(f), C.byref(c)))
        return int(f.value), int(c.value)

    def set_samples(self, coords, values):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = coords[:, None]
        n, dim = coords.shape
        values = np.ascontiguousarray(values, dtype=np.float64)
        self._keep = (coords, values)
        self._check(self._lib.gsm_multi_set_samples(self._h, _ptr(coords), int(dim), int(n), _ptr(values)))
        self.dim, self.n = int(dim), int(n)

    def set_samples_scaled(self, coords, values, scale_floor):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = coords[:, None]
        n, dim = coords.shape
        values = np.ascontiguousarray(values, dtype=np.float64)
        if values.shape != (n,):
            raise ValueError("values must have one entry per sample")
        self._keep = (coords, values)
        alpha, nan = C.c_double(0.0), C.c_int64(0)
        self._check(self._lib.gsm_multi_set_samples_scaled(self._h, _ptr(coords), int(dim), int(n), _ptr(values),
                                                           C.c_double(float(scale_floor)), C.byref(alpha), C.byref(nan)))
        self.dim, self.n = int(dim), int(n)
        return float(alpha.value), int(nan.value)

    def local_krige(self, vario, model, targets, maxneighbors, minneighbors=1, radius=0.0, want_var=True,
                    want_neighbors=False, out_mean=None, out_var=None, out_status=None):
        if want_neighbors:
            raise NotImplementedError("neighbour lists are a single-device test hook (Context.knn)")
        M = Context.target_count(targets)
        mean = _pool.empty(M, np.float64) if out_mean is None else out_mean
        var = (_pool.empty(M, np.float64) if out_var is None else out_var) if want_var else None
        status = _pool.empty(M, np.uint8) if out_status is None else out_status
        o = NeighborOpts(int(maxneighbors), int(minneighbors), float(radius))
        self._check(self._lib.gsm_multi_local_krige(self._h, C.byref(vario), C.byref(model), C.byref(targets), C.byref(o),
                                                    _ptr(mean), _ptr(var), _ptr(status)))
        return mean, var, status, None

    def idw(self, exponent, targets, maxneighbors=0, minneighbors=1, radius=0.0, out_mean=None):
        M = Context.target_count(targets)
        mean = _pool.empty(M, np.float64) if out_mean is None else out_mean
        status = _pool.empty(M, np.uint8)
        o = NeighborOpts(int(maxneighbors), int(minneighbors), float(radius)) if maxneighbors != 0 else None
        self._check(self._lib.gsm_multi_idw(self._h, float(exponent), C.byref(targets),
                                            C.byref(o) if o is not None else None, _ptr(mean), _ptr(status)))
        return mean, status

    def global_krige(self, vario, model, targets, want_var=True):
        M = Context.target_count(targets)
        mean = _pool.empty(M, np.float64)
        var = _pool.empty(M, np.float64) if want_var else None
        st = C.c_int32(0)
        self._check(self._lib.gsm_multi_global_krige(self._h, C.byref(vario), C.byref(model), C.byref(targets),
                                                     _ptr(mean), _ptr(var), C.byref(st)))
        return mean, var, bool(st.value)


class GlobalFit:
    def __init__(self, ctx: Context, handle):
        self.ctx, self._h = ctx, handle

    def status(self) -> bool:
        return bool(self.ctx._lib.gsm_global_krige_status(self._h))

    def predict(self, targets: Targets, want_var=True, out_mean=None, out_var=None):
        M = Context.target_count(targets)
        mean = _pool.empty(M, np.float64) if out_mean is None else out_mean
        var = (_pool.empty(M, np.float64) if out_var is None else out_var) if want_var else None
        self.ctx._check(self.ctx._lib.gsm_global_krige_predict(self._h, C.byref(targets), _ptr(mean), _ptr(var)))
        return mean, var

    def close(self):
        if self._h is not None:
            self.ctx._lib.gsm_global_krige_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_variogram(kind: int, sill: float, nugget: float, rng: float) -> Variogram:
    v = Variogram()
    v.kind, v.nstruct, v.sill, v.nugget, v.range = int(kind), 0, float(sill), float(nugget), float(rng)
    return v


def make_composite_variogram(structs) -> Variogram:
    """structs: iterable of (kind, sill, nugget, range, metric 3x3 or None)"""
    structs = list(structs)
    if not 1 <= len(structs) <= GSM_MAX_STRUCTURES:
        raise ValueError("a composite variogram takes 1..4 structures")
    v = Variogram()
    v.kind, v.nstruct, v.sill, v.nugget, v.range = 0, len(structs), 0.0, 0.0, 1.0
    for i, (kind, sill, nugget, rng, metric) in enumerate(structs):
        s = v.structs[i]
        s.kind, s.sill, s.nugget, s.range = int(kind), float(sill), float(nugget), float(rng)
        m = np.eye(3) if metric is None else np.asarray(metric, dtype=np.float64).reshape(3, 3)
        for e in range(9):
            s.metric[e] = float(m.flat[e])
    return v


def make_model(kind: int, mean: float = 0.0, exps=()) -> Model:
    m = Model()
    m.kind = int(kind)
    m.mean = float(mean)
    m.ndrift = len(exps)
    if len(exps) > GSM_MAX_DRIFT:
        raise ValueError("too many drift functions")
    for n, e in enumerate(exps):
        for d in range(GSM_MAX_DIM):
            m.exps[n][d] = int(e[d]) if d < len(e) else 0
    return m
