"""
c_oracle.py — ctypes front-end of oracle/_build/liblgssm_ref.so (the C restatement of the
reference's sequential recursions). TEST INFRASTRUCTURE ONLY — see lgssm_ref.c.

The descriptor mirrors include/tgp_b200.h's tgp_lgssm so tests hand identical inputs to the
oracle and to libtgpb200.so; this module deliberately has its own copy of the struct so it
does not import the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblgssm_ref.so")

_dp = C.POINTER(C.c_double)


class LGSSMDesc(C.Structure):
    _fields_ = [
        ("D", C.c_int32), ("M", C.c_int32), ("T", C.c_int64),
        ("ordering", C.c_int32), ("R_kind", C.c_int32),
        ("A", C.c_void_p), ("sA", C.c_int64),
        ("a", C.c_void_p), ("sa", C.c_int64),
        ("Q", C.c_void_p), ("sQ", C.c_int64),
        ("H", C.c_void_p), ("sH", C.c_int64),
        ("h", C.c_void_p), ("sh", C.c_int64),
        ("R", C.c_void_p), ("sR", C.c_int64),
        ("m0", C.c_void_p), ("P0", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("lgssm_ref.c", "lgssm_ref_steps.inc", "Makefile")
    ):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_pairwise_sum.restype = C.c_double
        _lib.oracle_pairwise_sum.argtypes = [C.c_void_p, C.c_int64]
    return _lib


def _per_step(x, inner_shape):
    """-> (contiguous float64 array kept alive, stride in elements). A leading-axis broadcast
    (NumPy stride 0) or an array without the T axis is passed as time-invariant (stride 0)."""
    x = np.asarray(x, dtype=np.float64)
    n_inner = int(np.prod(inner_shape)) if inner_shape else 1
    if x.ndim == len(inner_shape):
        return np.ascontiguousarray(x), 0
    if x.strides[0] == 0 or x.shape[0] == 1:
        return np.ascontiguousarray(x[0]), 0
    return np.ascontiguousarray(x), n_inner


def _colmajor(x):
    """(T?, r, c) mathematical matrices -> memory holding column-major r x c blocks."""
    x = np.asarray(x, dtype=np.float64)
    return np.swapaxes(x, -1, -2)


class Model:
    """Owns contiguous column-major copies of an LGSSM and the ctypes descriptor.

    Arguments are mathematical (row-major NumPy) arrays: As (T,D,D)|(D,D), as_ (T,D)|(D,),
    Qs like As, Hs (T,D)|(D,) for scalar emissions or (T,M,D)|(M,D), hs (T,)|() or (T,M)|(M,),
    Rs (T,)|() scalar emissions, (T,M)|(M,) diag, (T,M,M)|(M,M) dense."""

    def __init__(self, As, as_, Qs, m0, P0, Hs, hs, Rs, ordering="forward", T=None, M=1, R_kind=0):
        D = np.asarray(m0).shape[0]
        self.D, self.M = D, M
        self._keep = []
        d = LGSSMDesc()
        d.D, d.M = D, M
        d.ordering = 0 if ordering == "forward" else 1
        d.R_kind = R_kind

        def put(name, sname, arr, inner, colmajor=False):
            arr = np.asarray(arr, dtype=np.float64)
            if colmajor:
                arr = _colmajor(arr)
            a, s = _per_step(arr, inner)
            self._keep.append(a)
            setattr(d, name, a.ctypes.data)
            if sname:
                setattr(d, sname, s)
            return a

        A = put("A", "sA", As, (D, D), True)
        put("a", "sa", as_, (D,))
        put("Q", "sQ", Qs, (D, D), True)
        if M == 1 and R_kind == 0:
            put("H", "sH", Hs, (D,))
            put("h", "sh", hs, ())
            put("R", "sR", Rs, ())
        else:
            put("H", "sH", Hs, (D, M), True)  # after swap: memory (D, M) blocks == col-major M x D
            put("h", "sh", hs, (M,))
            rin = {0: (), 1: (M,), 2: (M, M)}[R_kind]
            put("R", "sR", Rs, rin, R_kind == 2)
        put("m0", None, m0, (D,))
        put("P0", None, P0, (D, D), True)
        if T is None:
            T = max(np.asarray(x).shape[0] if np.asarray(x).ndim > k else 1
                    for x, k in ((As, 2), (as_, 1), (Qs, 2)))
        d.T = int(T)
        self.T = int(T)
        self.desc = d

    @classmethod
    def from_lgssm(cls, m):
        """From an oracle.tgp_oracle.LGSSM."""
        if m.scalar:
            return cls(m.As, m.as_, m.Qs, m.m0, m.P0, m.Hs, m.hs, m.Rs, m.ordering, T=m.T)
        M = m.Hs.shape[1]
        return cls(m.As, m.as_, m.Qs, m.m0, m.P0, m.Hs, m.hs, m.Rs, m.ordering, T=m.T, M=M, R_kind=2)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def set_static(on: bool):
    lib().oracle_set_static(C.c_int(1 if on else 0))


def filter(model: Model, y, want_mp=True, want_steps=True):
    """-> dict(lml, lml_steps, m (T,D), P (T,D,D) mathematical)."""
    T, D = model.T, model.D
    y = np.ascontiguousarray(y, dtype=np.float64)
    m = np.empty((T, D)) if want_mp else None
    P = np.empty((T, D, D)) if want_mp else None
    steps = np.empty(T) if want_steps else None
    lml = C.c_double()
    fail = C.c_int64(-1)
    rc = lib().oracle_filter(C.byref(model.desc), _p(y), _p(m), C.c_int64(D), _p(P), C.c_int64(D * D),
                             C.byref(lml), _p(steps), C.byref(fail))
    if rc:
        raise RuntimeError(f"oracle_filter rc={rc} fail_t={fail.value}")
    return dict(lml=lml.value, lml_steps=steps, m=m, P=None if P is None else np.swapaxes(P, 1, 2))


def logpdf(model: Model, y):
    return filter(model, y, want_mp=False, want_steps=False)["lml"]


def logpdf_replicas(model: Model, ys):
    """`len(ys)` independent logpdf evaluations (rows of the C-contiguous 2-D array ys), one POSIX thread each.
    -> (lml per row, threads used)."""
    ys = np.ascontiguousarray(ys, dtype=np.float64)
    n, T = ys.shape
    assert T == model.T
    out = np.empty(n)
    nth = lib().oracle_logpdf_replicas(C.byref(model.desc), _p(ys), C.c_int64(T), C.c_int(n), _p(out))
    return out, int(nth)


def posterior(model: Model, y):
    T, D = model.T, model.D
    y = np.ascontiguousarray(y, dtype=np.float64)
    G = np.empty((T, D, D)); g = np.empty((T, D)); S = np.empty((T, D, D))
    mT = np.empty(D); PT = np.empty((D, D)); lml = C.c_double()
    rc = lib().oracle_posterior(C.byref(model.desc), _p(y), _p(G), _p(g), _p(S), _p(mT), _p(PT), C.byref(lml))
    if rc:
        raise RuntimeError(f"oracle_posterior rc={rc}")
    return dict(G=np.swapaxes(G, 1, 2), g=g, Sig=np.swapaxes(S, 1, 2), m_T=mT, P_T=PT.T, lml=lml.value)


def marginals(model: Model):
    T = model.T
    mean = np.empty(T); var = np.empty(T)
    rc = lib().oracle_marginals(C.byref(model.desc), _p(mean), _p(var))
    if rc:
        raise RuntimeError(f"oracle_marginals rc={rc}")
    return mean, var


def posterior_marginals(model: Model, y, R_new):
    T = model.T
    y = np.ascontiguousarray(y, dtype=np.float64)
    Rn = np.ascontiguousarray(np.atleast_1d(np.asarray(R_new, dtype=np.float64)))
    s = 0 if Rn.shape[0] == 1 else 1
    mean = np.empty(T); var = np.empty(T); lml = C.c_double()
    rc = lib().oracle_posterior_marginals(C.byref(model.desc), _p(y), _p(Rn), C.c_int64(s), _p(mean), _p(var),
                                          C.byref(lml))
    if rc:
        raise RuntimeError(f"oracle_posterior_marginals rc={rc}")
    return mean, var, lml.value
