"""ctypes loader of oracle/_build/libgsm_oracle.so (C twin of gsm_oracle.py).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import gsm_oracle as o

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OVario(C.Structure):
    _fields_ = [("kind", C.c_int), ("sill", C.c_double), ("nugget", C.c_double), ("range", C.c_double)]


class OModel(C.Structure):
    _fields_ = [("kind", C.c_int), ("mean", C.c_double), ("ndrift", C.c_int), ("exps", (C.c_int * 3) * 10)]


def load():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libgsm_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _LIB = C.CDLL(path)
        _LIB.gsmo_local.restype = C.c_int
    return _LIB


def _model(model):
    m = OModel()
    if isinstance(model, o.KrigingModel):
        m.kind, m.mean, m.ndrift = model.kind, model.mean, len(model.exps)
        for n, e in enumerate(model.exps):
            for d in range(3):
                m.exps[n][d] = e[d] if d < len(e) else 0
        v = OVario(model.vario.kind, model.vario.sill, model.vario.nugget, model.vario.range)
        return 1, m, v, 0.0
    return 2, m, OVario(1, 1.0, 0.0, 1.0), float(model.exponent)


def local(model, sX, z, sT, k, minn=1, radius=0.0, prob=True, use_tree=True, nthreads=1, want_nbr=False,
          knn_only=False):
    """Local prediction on ALREADY SCALED samples sX and explicit scaled targets sT (M x dim)."""
    lib = load()
    sX = np.ascontiguousarray(sX, dtype=np.float64)
    sT = np.ascontiguousarray(sT, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    n, dim = sX.shape
    M = sT.shape[0]
    k = n if (k > n or k < 1) else k
    minn = 1 if (minn > k or minn < 1) else minn
    mode, m, v, e = _model(model) if not knn_only else (0, OModel(), OVario(1, 1.0, 0.0, 1.0), 0.0)
    mean = np.full(M, np.nan)
    var = np.full(M, np.nan) if prob else None
    status = np.zeros(M, dtype=np.uint8)
    nbr = np.empty((M, k), dtype=np.int32) if (want_nbr or knn_only) else None
    cnt = np.empty(M, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = lib.gsmo_local(p(sX), C.c_int(dim), C.c_int(n), p(z), C.byref(v), C.byref(m), C.c_int(mode), C.c_double(e),
                        p(sT), C.c_long(M), C.c_int(k), C.c_int(minn), C.c_double(radius), C.c_int(int(use_tree)),
                        C.c_int(nthreads), p(mean), p(var), p(status), p(nbr), p(cnt))
    if rc != 0:
        raise RuntimeError("gsmo_local failed")
    return mean, var, status, nbr, cnt


def fitpredict(model, X, z, dom, *, prob=False, minneighbors=1, maxneighbors=10, neighborhood=None, first=0,
               count=None, use_tree=True, nthreads=1, want_nbr=False):
    """Neighbourhood fitpredict (models.jl:102-198) with the scaling of gsm_oracle.fitpredict."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    rng = model.vario.range if isinstance(model, o.KrigingModel) else np.inf
    alpha = o.scale_factor(rng, o.points_absmax(X), o._dom_absmax(dom))
    smodel = model.scaled(alpha) if isinstance(model, o.KrigingModel) else model
    sT = o._targets(dom.scaled(alpha), first, count) if isinstance(dom, o.Grid) else \
        o._targets(np.asarray(dom, dtype=np.float64) * alpha, first, count)
    radius = 0.0 if neighborhood is None else alpha * float(neighborhood)
    return local(smodel, X * alpha, z, sT, maxneighbors, minneighbors, radius, prob, use_tree, nthreads, want_nbr)
