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


def _scaled_problem(model, X, dom, neighborhood, first, count):
    """`_scale` of models.jl:272-284 applied to model, samples, targets and ball radius."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    rng = model.vario.range if isinstance(model, o.KrigingModel) else np.inf
    alpha = o.scale_factor(rng, o.points_absmax(X), o._dom_absmax(dom))
    smodel = model.scaled(alpha) if isinstance(model, o.KrigingModel) else model
    sT = o._targets(dom.scaled(alpha), first, count) if isinstance(dom, o.Grid) else \
        o._targets(np.asarray(dom, dtype=np.float64) * alpha, first, count)
    radius = 0.0 if neighborhood is None else alpha * float(neighborhood)
    return smodel, X * alpha, sT, radius


def fitpredict(model, X, z, dom, *, prob=False, minneighbors=1, maxneighbors=10, neighborhood=None, first=0,
               count=None, use_tree=True, nthreads=1, want_nbr=False):
    """Neighbourhood fitpredict (models.jl:102-198) with the scaling of gsm_oracle.fitpredict."""
    smodel, sX, sT, radius = _scaled_problem(model, X, dom, neighborhood, first, count)
    return local(smodel, sX, z, sT, maxneighbors, minneighbors, radius, prob, use_tree, nthreads, want_nbr)


def fitpredict_exact(model, X, z, dom, *, minneighbors=1, maxneighbors=10, neighborhood=None, first=0, count=None,
                     nthreads=1):
    """fitpredict twice on the same neighbour lists: the float64 reference algorithm and its binary128 evaluation.
    Returns (mu64, var64, status64, mu_q, var_q, ok_q)."""
    smodel, sX, sT, radius = _scaled_problem(model, X, dom, neighborhood, first, count)
    mu, var, st, nbr, cnt = local(smodel, sX, z, sT, maxneighbors, minneighbors, radius, True, True, nthreads, want_nbr=True)
    n = sX.shape[0]
    k = n if (maxneighbors > n or maxneighbors < 1) else maxneighbors
    minn = 1 if (minneighbors > k or minneighbors < 1) else minneighbors
    mu_q, var_q, ok_q, _ = local_quad(smodel, sX, z, sT, nbr, np.where(cnt >= minn, cnt, 0))
    return mu, var, st, mu_q, var_q, ok_q


# ----------------------------------------------------------------------------------------------
# quad-precision twin (oracle/gsm_oracle_q.c): the exact-arithmetic referee
# ----------------------------------------------------------------------------------------------
_QLIB = None


def load_quad():
    global _QLIB
    if _QLIB is None:
        path = os.path.join(_HERE, "_build", "libgsm_oracle_q.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _QLIB = C.CDLL(path)
        _QLIB.gsmq_local.restype = C.c_int
    return _QLIB


def local_quad(model, sX, z, sT, nbr, cnt):
    """The reference's Kriging equations for explicit neighbour lists, every operation in IEEE binary128.
    sX / sT: ALREADY SCALED samples / targets (float64 inputs, exactly what the float64 paths read); nbr: (M, k)
    sample rows, cnt: (M,) neighbours per target.  Returns (mean, var incl. +1e-10, ok, pivot ratio)."""
    lib = load_quad()
    sX = np.ascontiguousarray(sX, dtype=np.float64)
    sT = np.ascontiguousarray(sT, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    nbr = np.ascontiguousarray(nbr, dtype=np.int32)
    cnt = np.ascontiguousarray(cnt, dtype=np.int32)
    n, dim = sX.shape
    M, k = nbr.shape
    _, m, v, _ = _model(model)
    mean, var, cond = np.empty(M), np.empty(M), np.empty(M)
    ok = np.zeros(M, dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.gsmq_local(p(sX), C.c_int(dim), C.c_int(n), p(z), C.byref(v), C.byref(m), p(sT), C.c_long(M), C.c_int(k),
                        p(nbr), p(cnt), p(mean), p(var), p(ok), p(cond))
    if rc != 0:
        raise RuntimeError("gsmq_local failed")
    return mean, var, ok.astype(bool), cond


def grid_subset_exact(model, X, z, grid, lin_idx, maxneighbors, nthreads=8):
    """Strided subset `lin_idx` of a full `fitpredict(model, data, grid)`: scaling with the alpha of the WHOLE grid
    (models.jl:272-284), neighbour lists from the C oracle's k-NN, then both referees on the same lists:
    the float64 reference algorithm (Bunch-Kaufman, raw monomials) and its binary128 evaluation.
    Returns dict(mu64, var64, mu_q, var_q, ok_q, nbr, cnt, pivot_ratio)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    alpha = o.scale_factor(model.vario.range, o.points_absmax(X), grid.bbox_absmax())
    smodel = model.scaled(alpha)
    sg = grid.scaled(alpha)
    sT = np.stack([sg.centroids(int(i), 1)[0] for i in lin_idx])
    sX = X * alpha
    mu64, var64, st, nbr, cnt = local(smodel, sX, z, sT, maxneighbors, 1, 0.0, True, True, nthreads, want_nbr=True)
    mu_q, var_q, ok_q, piv = local_quad(smodel, sX, z, sT, nbr, cnt)
    return dict(mu64=mu64, var64=var64, st64=st, mu_q=mu_q, var_q=var_q, ok_q=ok_q, nbr=nbr, cnt=cnt, pivot_ratio=piv)
