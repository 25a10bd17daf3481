"""
_lib.py — ctypes binding of libtgpb200.so (include/tgp_b200.h). This is the Python stand-in for the
Julia `ccall` glue: pure marshalling, no arithmetic. There is no CPU fallback — if
the shared library is missing, or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtgpb200.so")

TGP_OK, TGP_EINVAL, TGP_ENOTPD, TGP_ECUDA, TGP_ENOMEM, TGP_EUNSUPPORTED = range(6)
TGP_FORWARD, TGP_REVERSE = 0, 1
TGP_R_SCALAR, TGP_R_DIAG, TGP_R_DENSE = 0, 1, 2
TGP_OPT_ALGO, TGP_OPT_CHUNK, TGP_OPT_SS_TOL, TGP_OPT_TIMING, TGP_OPT_SS_PREFIX = 1, 2, 3, 4, 5
TGP_ALGO_AUTO, TGP_ALGO_SCAN = 0, 1
TGP_OPT_DENSE_MATH = 6
TGP_OPT_DEFER_STATUS = 7
TGP_OPT_SHARD_OVERLAP = 8
TGP_SHARD_HALO = 3072
TGP_DENSE_F64, TGP_DENSE_TF32X3 = 0, 1


class TGPError(RuntimeError):
    """Non-zero status from libtgpb200 (the reference throws ErrorException, lgssm.jl:202-208)."""

    def __init__(self, code, msg):
        super().__init__(f"libtgpb200 error {code}: {msg}")
        self.code = code


class DimensionMismatch(TGPError, ValueError):
    pass


class PosDefException(TGPError, ArithmeticError):
    """A covariance that must be factorised is not positive definite (Julia: PosDefException)."""


class tgp_lgssm(C.Structure):
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


_SIGS = {
    "tgp_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "tgp_destroy": (None, [C.c_void_p]),
    "tgp_last_error": (C.c_char_p, [C.c_void_p]),
    "tgp_version": (C.c_char_p, []),
    "tgp_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "tgp_get_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "tgp_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tgp_get_timing": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "tgp_logpdf": (C.c_int, [C.c_void_p, C.POINTER(tgp_lgssm), C.c_void_p, C.c_void_p, C.c_void_p]),
    "tgp_filter": (C.c_int, [C.c_void_p, C.POINTER(tgp_lgssm), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                             C.c_void_p]),
    "tgp_posterior": (C.c_int, [C.c_void_p, C.POINTER(tgp_lgssm), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "tgp_marginals": (C.c_int, [C.c_void_p, C.POINTER(tgp_lgssm), C.c_void_p, C.c_void_p]),
    "tgp_marginals_diag": (C.c_int, [C.c_void_p, C.POINTER(tgp_lgssm), C.c_void_p, C.c_void_p]),
    "tgp_lti_components": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "tgp_posterior_marginals": (C.c_int, [C.c_void_p, C.POINTER(tgp_lgssm), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "tgp_debug_tc_gemm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "tgp_elem_size": (C.c_int, [C.c_int]),
    "tgp_shard_reduce": (C.c_int, [C.c_void_p, C.POINTER(tgp_lgssm), C.c_void_p, C.c_void_p]),
    "tgp_shard_xchg_size": (C.c_int, [C.c_int]),
    "tgp_shard_phase1": (C.c_int, [C.c_void_p, C.POINTER(tgp_lgssm), C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "tgp_shard_phase2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "tgp_synchronize": (C.c_int, [C.c_void_p]),
    "tgp_shard_logpdf": (C.c_int, [C.c_void_p, C.POINTER(tgp_lgssm), C.c_void_p, C.c_int, C.c_int]),
    "tgp_shard_result": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tgp_shard_partial": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tgp_xchg_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tgp_xchg_open": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tgp_shard_prefix": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}
EXPORTS = tuple(_SIGS)

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load libtgpb200.so (built in-tree by build.py). Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `python temporalgps.jl_b200/build.py` "
                    "(nvcc, sm_100a). There is no CPU fallback for the LGSSM hot path.")
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGS.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def ptr(x):
    """Raw address of a NumPy array, a torch tensor (host or CUDA), an int, or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        if x.dtype != np.float64 or not x.flags.c_contiguous:
            raise TypeError("arrays handed to libtgpb200 must be C-contiguous float64")
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        import torch
        if x.dtype != torch.float64 or not x.is_contiguous():
            raise TypeError("tensors handed to libtgpb200 must be contiguous float64")
        return x.data_ptr()
    raise TypeError(f"cannot take the address of {type(x)!r}")


class Handle:
    """One tgp_handle = one GPU + stream + workspace arena."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        L = lib()
        rc = L.tgp_create(C.byref(self._h), int(device))
        if rc != TGP_OK:
            msg = L.tgp_last_error(None).decode()
            self._h = None
            raise TGPError(rc, msg or "tgp_create failed")
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.tgp_destroy(self._h)
            self._h = None

    __del__ = close

    def check(self, rc):
        if rc == TGP_OK:
            return
        msg = lib().tgp_last_error(self._h).decode()
        if rc == TGP_ENOTPD:
            raise PosDefException(rc, msg)
        if rc == TGP_EINVAL and "Dimension mismatch" in msg:
            raise DimensionMismatch(rc, msg)
        raise TGPError(rc, msg)

    def set_option(self, opt, value):
        self.check(lib().tgp_set_option(self._h, int(opt), int(value)))

    def set_algo(self, algo):
        self.set_option(TGP_OPT_ALGO, algo)

    def set_chunk(self, chunk):
        self.set_option(TGP_OPT_CHUNK, chunk)

    def set_ss_tol(self, tol):
        self.set_option(TGP_OPT_SS_TOL, int(np.float64(tol).view(np.int64)))

    def set_dense_math(self, mode):
        self.set_option(TGP_OPT_DENSE_MATH, mode)

    def tc_gemm(self, X, Y, symmetric=False):
        """Test hook: X' Y through the tcgen05 3xTF32 kernel. X (K, Mx), Y (K, N) float32; returns (Mx, N) float32."""
        Xf = np.asfortranarray(X, dtype=np.float32)
        Yf = np.asfortranarray(Y, dtype=np.float32)
        K, Mx = Xf.shape
        N = Yf.shape[1]
        Cm = np.zeros((Mx, N), dtype=np.float32, order="F")
        self.check(lib().tgp_debug_tc_gemm(self._h, K, Mx, N, Xf.ctypes.data, Yf.ctypes.data, Cm.ctypes.data, 1 if symmetric else 0))
        return Cm

    def set_timing(self, on: bool):
        self.set_option(TGP_OPT_TIMING, 1 if on else 0)

    def timing(self, cap: int = 32):
        """-> list of (kernel name, total ms, launches), sorted by total device time."""
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        n = (C.c_int64 * cap)()
        k = lib().tgp_get_timing(self._h, cap, names, ms, n)
        return [(names[i].decode(), ms[i], n[i]) for i in range(min(k, cap))]

    def set_stream(self, cuda_stream: int):
        self.check(lib().tgp_set_stream(self._h, cuda_stream))

    def counters(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.check(lib().tgp_get_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(launches=a.value, h2d_bytes=b.value, d2h_bytes=c.value)

    # raw entry points -------------------------------------------------------------------------
    def logpdf(self, desc, y, lml_out, lml_per_step=None):
        self.check(lib().tgp_logpdf(self._h, C.byref(desc), ptr(y), ptr(lml_out), ptr(lml_per_step)))

    def filter(self, desc, y, m_f, s_m, P_f, s_P, lml_out=None):
        self.check(lib().tgp_filter(self._h, C.byref(desc), ptr(y), ptr(m_f), int(s_m), ptr(P_f), int(s_P), ptr(lml_out)))

    def posterior(self, desc, y, G, g, Sig, m_T, P_T):
        self.check(lib().tgp_posterior(self._h, C.byref(desc), ptr(y), ptr(G), ptr(g), ptr(Sig), ptr(m_T), ptr(P_T)))

    def marginals(self, desc, mean_out, cov_out):
        self.check(lib().tgp_marginals(self._h, C.byref(desc), ptr(mean_out), ptr(cov_out)))

    def marginals_diag(self, desc, mean_out, var_out):
        self.check(lib().tgp_marginals_diag(self._h, C.byref(desc), ptr(mean_out), ptr(var_out)))

    def lti_components(self, F, P, t, A_out, Q_out, F0=None):
        """F, P, F0: (D, D) host arrays (mathematical orientation); t, A_out, Q_out: host arrays or device tensors."""
        F = np.asfortranarray(F, dtype=np.float64)
        P = np.asfortranarray(P, dtype=np.float64)
        F0 = None if F0 is None else np.asfortranarray(F0, dtype=np.float64)
        D = F.shape[0]
        T = t.numel() if hasattr(t, "numel") else len(t)
        self.check(lib().tgp_lti_components(self._h, D, int(T), F.ctypes.data, None if F0 is None else F0.ctypes.data, P.ctypes.data,
                                            ptr(t), ptr(A_out), ptr(Q_out)))

    def posterior_marginals(self, desc, y, R_new, sRnew, mean_out, var_out, lml_out=None):
        self.check(lib().tgp_posterior_marginals(self._h, C.byref(desc), ptr(y), ptr(R_new), int(sRnew), ptr(mean_out),
                                                 ptr(var_out), ptr(lml_out)))

    def shard_reduce(self, desc, y, elem_out):
        self.check(lib().tgp_shard_reduce(self._h, C.byref(desc), ptr(y), ptr(elem_out)))

    def shard_prefix(self, D, elems, m0, P0):
        return shard_prefix(D, elems, m0, P0, self)

    def shard_phase1(self, desc, y, rank, world, xchg_out):
        self.check(lib().tgp_shard_phase1(self._h, C.byref(desc), ptr(y), int(rank), int(world), ptr(xchg_out)))

    def shard_logpdf(self, desc, y, rank, world):
        """One-launch sharded logpdf (enqueued, not synchronised). Raises TGPError(TGP_EUNSUPPORTED) if the model is outside its range."""
        self.check(lib().tgp_shard_logpdf(self._h, C.byref(desc), ptr(y), int(rank), int(world)))

    def shard_result(self, lml_total):
        self.check(lib().tgp_shard_result(self._h, ptr(lml_total)))

    def shard_partial(self, lml_shard):
        self.check(lib().tgp_shard_partial(self._h, ptr(lml_shard)))

    def synchronize(self):
        self.check(lib().tgp_synchronize(self._h))

    def xchg_create(self, rank, world, slot_doubles):
        """-> the 64-byte CUDA IPC handle of this rank's exchange buffer."""
        buf = (C.c_ubyte * 64)()
        self.check(lib().tgp_xchg_create(self._h, int(rank), int(world), int(slot_doubles), buf))
        return bytes(buf)

    def xchg_open(self, handles_all: bytes):
        self.check(lib().tgp_xchg_open(self._h, handles_all))

    def shard_phase2(self, xchg_all, lml_partial):
        self.check(lib().tgp_shard_phase2(self._h, ptr(xchg_all), ptr(lml_partial)))


def shard_prefix(D, elems, m0, P0, handle=None):
    """tgp_shard_prefix: fold scan elements 0..n-1 into (m0, P0). Host arithmetic; works without a device."""
    n = 0 if elems is None else int(np.asarray(elems).shape[0])
    m_in = np.empty(D)
    P_in = np.empty((D, D))
    e = None if n == 0 else np.ascontiguousarray(elems, dtype=np.float64)
    rc = lib().tgp_shard_prefix(handle._h if handle is not None else None, int(D), n, ptr(e),
                                ptr(np.ascontiguousarray(m0, dtype=np.float64)), ptr(np.ascontiguousarray(P0, dtype=np.float64)),
                                ptr(m_in), ptr(P_in))
    if rc != TGP_OK:
        raise TGPError(rc, "tgp_shard_prefix failed")
    return m_in, P_in


_default = {}


def default_handle(device: int = 0) -> Handle:
    h = _default.get(device)
    if h is None:
        h = _default[device] = Handle(device)
    return h
