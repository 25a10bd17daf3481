"""
lgssm.py — host-side mirror of the reference's L2 interface (src/models/lgssm.jl,
gauss_markov_model.jl, missings.jl) over the C ABI of libtgpb200.so.

Same names and argument meaning as the reference: `LGSSM`, `GaussMarkovModel`, `Forward`/`Reverse`,
`Gaussian`, `logpdf(model, y)` (lgssm.jl:147), `_filter` (:171), `posterior` (:193), `marginals`
(:99), plus the missing-data wrappers of missings.jl:8-23. Like the reference, the missing-data
path is selected by the TYPE of y — `numpy.ma.MaskedArray` plays `Vector{Union{Missing,T}}` (masked
== missing); a plain float array is never scanned for NaN. All arithmetic of
the recursions happens in the CUDA library; this module only lays arrays out the way Julia would
(column-major blocks, stride 0 for `Fill`) and maps status codes to exceptions.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import (TGP_DENSE_F64, TGP_DENSE_TF32X3, TGP_OPT_DENSE_MATH, DimensionMismatch, Handle, PosDefException, TGPError,
                   default_handle, tgp_lgssm)

Forward = "forward"   # gauss_markov_model.jl:1
Reverse = "reverse"   # gauss_markov_model.jl:3

LARGE_VAR = 1e15      # _large_var_const(), missings.jl:43


class Fill:
    """A time-invariant per-step array (Julia `Fill`, lti_sde.jl:148-160): one value, length T."""

    def __init__(self, value, T: int):
        self.value = np.asarray(value, dtype=np.float64)
        self.T = int(T)

    def __len__(self):
        return self.T

    def __getitem__(self, t):
        return self.value

    @property
    def shape(self):
        return (self.T,) + self.value.shape


class DeviceSteps:
    """A per-step array that already lives on the GPU in the library's layout (contiguous per step, matrices column-major), e.g.
    the transitions written by tgp_lti_components. It crosses the boundary by address; nothing is copied or re-laid out."""

    def __init__(self, tensor, T: int, inner_shape):
        self.tensor, self.T, self.inner_shape = tensor, int(T), tuple(inner_shape)
        if tensor.numel() != self.T * int(np.prod(self.inner_shape)):
            raise DimensionMismatch(_lib.TGP_EINVAL, "Dimension mismatch. device per-step array has the wrong number of elements")

    def __len__(self):
        return self.T

    @property
    def shape(self):
        return (self.T,) + self.inner_shape

    def numpy(self):
        """Host copy in mathematical orientation (T, *inner_shape) — for inspection and tests."""
        a = self.tensor.detach().cpu().numpy().reshape((self.T,) + self.inner_shape[::-1])
        return np.swapaxes(a, -1, -2) if len(self.inner_shape) == 2 else a


class _DevAddr:
    """What _Marshalled keeps alive for a DeviceSteps field (mimics the `.ctypes.data` of a host array)."""

    def __init__(self, tensor):
        self.tensor = tensor
        self.ctypes = type("addr", (), {"data": tensor.data_ptr()})()


@dataclass
class Gaussian:
    """gaussian.jl:16-19."""
    m: np.ndarray
    P: np.ndarray


@dataclass
class GaussMarkovModel:
    """gauss_markov_model.jl:20-32 — x[t] = A[t] x[t-1] + a[t] + N(0, Q[t])."""
    ordering: str
    As: object   # Fill | (T, D, D)
    as_: object  # Fill | (T, D)
    Qs: object   # Fill | (T, D, D)
    x0: Gaussian

    def __len__(self):
        return len(self.As)


@dataclass
class ScalarEmissions:
    """StructArray{ScalarOutputLGC} (lti_sde.jl:88-101): y[t] = H[t]·x[t] + h[t] + N(0, R[t])."""
    Hs: object   # Fill | (T, D)
    hs: object   # Fill | (T,)
    Rs: object   # Fill | (T,)


@dataclass
class SmallOutputEmissions:
    """StructArray{SmallOutputLGC} (lti_sde.jl:103-109): y[t] = H[t] x[t] + h[t] + N(0, R[t]), y[t] in R^M.
    Rs: (T, M) | Fill((M,)) = diagonal (the space-time path's `Diagonal`), or (T, M, M) | Fill((M, M)) = dense."""
    Hs: object   # Fill | (T, M, D)
    hs: object   # Fill | (T, M)
    Rs: object

    @property
    def M(self):
        H = self.Hs.value if isinstance(self.Hs, Fill) else np.asarray(self.Hs)[0]
        return int(np.asarray(H).shape[0])

    @property
    def r_dense(self):
        R = self.Rs.value if isinstance(self.Rs, Fill) else np.asarray(self.Rs)[0]
        return np.ndim(R) == 2


@dataclass
class LargeOutputEmissions(SmallOutputEmissions):
    """StructArray{LargeOutputLGC} (linear_gaussian_conditionals.jl:153-204): the same conditional y | x ~ N(H x + h, R) as
    SmallOutputLGC, which the reference evaluates by another route (Cholesky in the latent space, jitter 1e-10) when Dobs > Dlat.
    The library has ONE route for vector observations, so this marshals exactly like SmallOutputEmissions; results agree with the
    reference's to its own jitter (~1e-10 relative, inside the 1e-6 / 1e-5 parity band)."""


@dataclass
class BottleneckEmissions:
    """StructArray{BottleneckLGC} (linear_gaussian_conditionals.jl:258-335): y | x ~ N(A (H x + h) + a, Q) with a low-dimensional
    projection (H, h) and a fan-out LargeOutputLGC (A, a, Q). Marshalled as the equivalent single conditional
    N((A H) x + (A h + a), Q); the reference's 1e-12 jitters on the projected covariance are below the parity band."""
    Hs: object        # Fill | (T, K, D)
    hs: object        # Fill | (T, K)
    fan_out: LargeOutputEmissions   # A: (T, M, K), a: (T, M), Q

    def collapse(self, T) -> SmallOutputEmissions:
        f = self.fan_out
        fills = all(isinstance(v, Fill) for v in (self.Hs, self.hs, f.Hs, f.hs))
        n = 1 if fills else T
        H = _dense_steps(self.Hs, n)
        h = _dense_steps(self.hs, n)
        A = _dense_steps(f.Hs, n)
        a = _dense_steps(f.hs, n)
        Hs = np.einsum("tmk,tkd->tmd", A, H)
        hs = np.einsum("tmk,tk->tm", A, h) + a
        if fills:
            return SmallOutputEmissions(Fill(Hs[0], T), Fill(hs[0], T), f.Rs)
        return SmallOutputEmissions(Hs, hs, f.Rs)

    @property
    def M(self):
        return self.fan_out.M

    @property
    def r_dense(self):
        return self.fan_out.r_dense


def _dense_steps(v, n):
    if isinstance(v, Fill):
        return np.broadcast_to(v.value, (n,) + v.value.shape)
    v = np.asarray(v, dtype=np.float64)
    return v if v.shape[0] == n else np.broadcast_to(v[0], (n,) + v.shape[1:])


@dataclass
class LGSSM:
    """lgssm.jl:9-12."""
    transitions: GaussMarkovModel
    emissions: object   # ScalarEmissions | SmallOutputEmissions

    def __len__(self):
        return len(self.transitions)

    @property
    def ordering(self):
        return self.transitions.ordering

    @property
    def D(self):
        return int(np.asarray(self.transitions.x0.m).shape[0])


def _per_step(x, T, inner_ndim, colmajor=False):
    """-> (C-contiguous float64 array, stride in elements). Fill / un-batched arrays get stride 0."""
    if isinstance(x, Fill):
        v = x.value
        if colmajor:
            v = v.T
        return np.ascontiguousarray(v, dtype=np.float64), 0
    if isinstance(x, DeviceSteps):
        if len(x) != T or len(x.inner_shape) != inner_ndim:
            raise DimensionMismatch(_lib.TGP_EINVAL, f"Dimension mismatch. per-step array has shape {x.shape}, model has T={T}")
        return _DevAddr(x.tensor), int(np.prod(x.inner_shape))
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == inner_ndim:
        return np.ascontiguousarray(x.T if colmajor else x), 0
    if x.shape[0] != T:
        raise DimensionMismatch(_lib.TGP_EINVAL, f"Dimension mismatch. per-step array has length {x.shape[0]}, model has {T}")
    if x.strides[0] == 0:
        v = x[0]
        return np.ascontiguousarray(v.T if colmajor else v), 0
    if colmajor:
        x = np.swapaxes(x, -1, -2)
    x = np.ascontiguousarray(x)
    return x, int(np.prod(x.shape[1:])) if x.ndim > 1 else 1


class _Marshalled:
    """Owns the contiguous column-major copies referenced by a tgp_lgssm descriptor."""

    def __init__(self, model: LGSSM):
        tr, em = model.transitions, model.emissions
        T = len(model)
        D = model.D
        if isinstance(em, BottleneckEmissions):
            em = em.collapse(T)
        d = tgp_lgssm()
        vector = isinstance(em, SmallOutputEmissions)
        M = em.M if vector else 1
        d.D, d.M, d.T = D, M, T
        d.ordering = _lib.TGP_FORWARD if tr.ordering == Forward else _lib.TGP_REVERSE
        d.R_kind = (_lib.TGP_R_DENSE if em.r_dense else _lib.TGP_R_DIAG) if vector else _lib.TGP_R_SCALAR
        self.keep = []
        if vector:   # H is M x D column-major per step; y is (T, M) with M fastest
            emis = (("H", "sH", em.Hs, 2, True), ("h", "sh", em.hs, 1, False), ("R", "sR", em.Rs, 2 if em.r_dense else 1, em.r_dense))
        else:
            emis = (("H", "sH", em.Hs, 1, False), ("h", "sh", em.hs, 0, False), ("R", "sR", em.Rs, 0, False))
        for name, sname, arr, nd, cm in (("A", "sA", tr.As, 2, True), ("a", "sa", tr.as_, 1, False), ("Q", "sQ", tr.Qs, 2, True)) + emis:
            a, s = _per_step(arr, T, nd, cm)
            self.keep.append(a)
            setattr(d, name, a.ctypes.data)
            setattr(d, sname, s)
        m0 = np.ascontiguousarray(tr.x0.m, dtype=np.float64)
        P0 = np.ascontiguousarray(np.asarray(tr.x0.P, dtype=np.float64).T)
        if m0.shape != (D,) or P0.shape != (D, D):
            raise DimensionMismatch(_lib.TGP_EINVAL, "Dimension mismatch. x0 has the wrong shape")
        self.keep += [m0, P0]
        d.m0, d.P0 = m0.ctypes.data, P0.ctypes.data
        self.desc = d
        self.T, self.D, self.M = T, D, M


def _check_inputs(model: LGSSM, y):
    """lgssm.jl:202-208."""
    if len(model) != len(y):
        raise DimensionMismatch(_lib.TGP_EINVAL, f"Dimension mismatch. length(prior) is {len(model)}, but length(y) is {len(y)}")


def _host_y(y):
    """Observations as the library takes them: a C-contiguous float64 host array, or a device-resident
    torch tensor passed through by address (the library detects device pointers)."""
    if hasattr(y, "data_ptr"):
        return y
    return np.ascontiguousarray(y, dtype=np.float64)


# ---- missing data (missings.jl:25-74, LGC:143-151): stays on the host, kernels see plain R_t ------------------
def transform_model_and_obs(model: LGSSM, y):
    """-> (model with R := 1e15 where y is missing, y with 0 there, number of missing observation DIMENSIONS).
    Scalar observations: masked == missing. Vector observations (T, M): a fully masked row is a missing observation (any R:
    R_t := 1e15 I, missings.jl:59-74, 97); a partially masked row needs diagonal R (R_t[i] := 1e15 for the masked entries,
    LGC:143-151 with missings.jl:76-79 — the reference raises a MethodError for dense R there)."""
    miss = np.ma.getmaskarray(y)
    y = np.array(np.ma.getdata(y), dtype=np.float64, copy=True)
    T = len(model)
    if isinstance(model.emissions, BottleneckEmissions):
        model = replace(model, emissions=model.emissions.collapse(T))
    em = model.emissions
    Rs = em.Rs
    if not isinstance(em, SmallOutputEmissions):
        Rs = np.full(T, float(Rs.value)) if isinstance(Rs, Fill) else np.array(Rs, dtype=np.float64, copy=True)
        Rs[miss] = LARGE_VAR
        y[miss] = 0.0
        return replace(model, emissions=replace(em, Rs=Rs)), y, int(miss.sum())
    M = em.M
    if miss.shape != (T, M):
        raise DimensionMismatch(_lib.TGP_EINVAL, f"Dimension mismatch. y has shape {miss.shape}, the model emits {(T, M)}")
    whole = miss.all(axis=1)
    partial = miss.any(axis=1) & ~whole
    dense = em.r_dense
    if partial.any() and dense:
        raise TGPError(_lib.TGP_EUNSUPPORTED, "element-wise missing observations need a Diagonal observation covariance "
                                              "(linear_gaussian_conditionals.jl:143-151)")
    Rv = Rs.value if isinstance(Rs, Fill) else None
    if dense:
        Rs = np.broadcast_to(Rv, (T, M, M)).copy() if Rv is not None else np.array(Rs, dtype=np.float64, copy=True)
        Rs[whole] = LARGE_VAR * np.eye(M)
    else:
        Rs = np.broadcast_to(Rv, (T, M)).copy() if Rv is not None else np.array(Rs, dtype=np.float64, copy=True)
        Rs[miss] = LARGE_VAR
    y[miss] = 0.0
    return replace(model, emissions=replace(em, Rs=Rs)), y, int(miss.sum())


def _maybe_missing(model, y):
    if isinstance(y, np.ma.MaskedArray):          # dispatch on type, as missings.jl:8-23 does
        return transform_model_and_obs(model, y)
    return model, y, 0


def replace_observation_noise_cov(model: LGSSM, Rs_new) -> LGSSM:
    """missings.jl:35-41."""
    em = model.emissions
    if isinstance(em, BottleneckEmissions):
        return replace(model, emissions=replace(em, fan_out=replace(em.fan_out, Rs=Rs_new)))
    return replace(model, emissions=replace(em, Rs=Rs_new))


# ---- the five entry points -----------------------------------------------------------------------
def logpdf(model: LGSSM, y, handle: Optional[Handle] = None, per_step: bool = False):
    """logpdf(model::LGSSM, y) — lgssm.jl:147-151 (+ missings.jl:8-13 when y has NaN)."""
    _check_inputs(model, y)
    h = handle or default_handle()
    model, y, n_missing = _maybe_missing(model, y)
    mm = _Marshalled(model)
    out = np.zeros(1)
    steps = np.empty(mm.T) if per_step else None
    h.logpdf(mm.desc, _host_y(y), out, steps)
    lml = float(out[0]) + n_missing * math.log(2.0 * math.pi * LARGE_VAR) / 2.0
    return (lml, steps) if per_step else lml


def _filter(model: LGSSM, y, handle: Optional[Handle] = None):
    """_filter(model, y) — lgssm.jl:171-187. -> (ms (T, D), Ps (T, D, D)) filtering distributions."""
    _check_inputs(model, y)
    h = handle or default_handle()
    model, y, _ = _maybe_missing(model, y)
    mm = _Marshalled(model)
    T, D = mm.T, mm.D
    ms = np.empty((T, D))
    Ps = np.empty((T, D, D))
    h.filter(mm.desc, _host_y(y), ms, D, Ps, D * D, None)
    return ms, np.swapaxes(Ps, 1, 2)


def posterior(model: LGSSM, y, handle: Optional[Handle] = None) -> LGSSM:
    """posterior(prior::LGSSM, y) — lgssm.jl:193-200: the Reverse-ordered posterior LGSSM."""
    _check_inputs(model, y)
    h = handle or default_handle()
    model, y, _ = _maybe_missing(model, y)
    mm = _Marshalled(model)
    T, D = mm.T, mm.D
    G = np.empty((T, D, D)); g = np.empty((T, D)); S = np.empty((T, D, D))
    mT = np.empty(D); PT = np.empty((D, D))
    h.posterior(mm.desc, _host_y(y), G, g, S, mT, PT)
    new_order = Reverse if model.ordering == Forward else Forward
    tr = GaussMarkovModel(new_order, np.swapaxes(G, 1, 2), g, np.swapaxes(S, 1, 2), Gaussian(mT, PT.T))
    return LGSSM(tr, model.emissions)


def _emission_dim(model: LGSSM) -> int:
    return model.emissions.M if isinstance(model.emissions, (SmallOutputEmissions, BottleneckEmissions)) else 0


def marginals(model: LGSSM, handle: Optional[Handle] = None) -> Tuple[np.ndarray, np.ndarray]:
    """marginals(model::LGSSM) — lgssm.jl:99-115. Scalar emissions -> (means (T,), variances (T,)); vector emissions ->
    (means (T, M), covariances (T, M, M))."""
    h = handle or default_handle()
    mm = _Marshalled(model)
    M = _emission_dim(model)
    if M == 0:
        mean = np.empty(mm.T)
        var = np.empty(mm.T)
        h.marginals(mm.desc, mean, var)
        return mean, var
    mean = np.empty((mm.T, M))
    cov = np.empty((mm.T, M, M))
    h.marginals(mm.desc, mean, cov)
    return mean, np.swapaxes(cov, 1, 2)


def marginals_diag(model: LGSSM, handle: Optional[Handle] = None) -> Tuple[np.ndarray, np.ndarray]:
    """marginals_diag(model::LGSSM) — lgssm.jl:125-141 (predict_marginals, LGC:63-68): means and the DIAGONAL of the emission-space
    covariances, (T,) / (T, M)."""
    h = handle or default_handle()
    mm = _Marshalled(model)
    M = _emission_dim(model)
    shape = (mm.T,) if M == 0 else (mm.T, M)
    mean = np.empty(shape)
    var = np.empty(shape)
    h.marginals_diag(mm.desc, mean, var)
    return mean, var


def posterior_marginals(model: LGSSM, y, Rs_new, handle: Optional[Handle] = None, return_lml: bool = False):
    """marginals_diag(replace_observation_noise_cov(posterior(model, y), Rs_new)) — the chain of posterior_lti_sde.jl:27-36 in one
    library call (scalar observations, Forward: fused, (G, g, Sigma) never materialised)."""
    _check_inputs(model, y)
    h = handle or default_handle()
    model, y, n_missing = _maybe_missing(model, y)
    mm = _Marshalled(model)
    M = _emission_dim(model)
    if M == 0:
        Rn, sRn = _per_step(Rs_new if isinstance(Rs_new, Fill) or np.ndim(Rs_new) else Fill(Rs_new, mm.T), mm.T, 0)
        Rn = np.atleast_1d(Rn)
        shape = (mm.T,)
    else:
        dense = model.emissions.r_dense
        Rn, sRn = _per_step(Rs_new, mm.T, 2 if dense else 1, dense)
        shape = (mm.T, M)
    mean = np.empty(shape)
    var = np.empty(shape)
    lml = np.zeros(1)
    h.posterior_marginals(mm.desc, _host_y(y), Rn, sRn, mean, var, lml)
    if return_lml:
        return mean, var, float(lml[0]) + n_missing * math.log(2.0 * math.pi * LARGE_VAR) / 2.0
    return mean, var
