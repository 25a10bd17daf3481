"""
gp.py — host-side mirror of the reference's GP layer (src/gp/lti_sde.jl, posterior_lti_sde.jl):
`GP`, `to_sde`, `f(x, σ²)`, `logpdf`, `posterior`, `marginals`, `mean_and_var`. It is the CALLER of
the hot path: O(1) (regular spacing) or O(T) (irregular) model set-up on the host with NumPy, then
every recursion runs in libtgpb200.so through lgssm.py. The storage tag `B200Storage` plays the
role of the reference's `StorageType` plug-in point (storage_types.jl:1, lti_sde.jl:12-14).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Union

import numpy as np
from scipy.linalg import expm
from scipy.special import ive

from . import lgssm as L
from .lgssm import Fill, Forward, Gaussian, GaussMarkovModel, LGSSM, ScalarEmissions, SmallOutputEmissions


# ---- storage tags (storage_types.jl) -------------------------------------------------------------
@dataclass(frozen=True)
class B200Storage:
    """Selects the B200 library for the LGSSM recursions. The small-state scan kernels are FP64 whatever `dtype` says;
    on the large-state / vector-observation path float32 selects the FP32-storage tcgen05 kernels (3xTF32 products,
    TGP_DENSE_TF32X3), float64 the FP64 step."""
    device: int = 0
    dtype: type = np.float64

    def handle(self):
        h = L.default_handle(self.device)
        h.set_option(L.TGP_OPT_DENSE_MATH, L.TGP_DENSE_TF32X3 if np.dtype(self.dtype) == np.float32 else L.TGP_DENSE_F64)
        return h


def SArrayStorage(dtype=np.float64):   # accepted for source compatibility; same path
    return B200Storage(0, dtype)


def ArrayStorage(dtype=np.float64):
    return B200Storage(0, dtype)


# ---- inputs --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class RegularSpacing:
    """regular_data.jl:8-22 — t0, t0 + Δt, ..., N points."""
    t0: float
    dt: float
    N: int

    def __len__(self):
        return self.N

    def collect(self):
        return self.t0 + np.arange(self.N, dtype=np.float64) * self.dt


def _times(x):
    return x.collect() if isinstance(x, RegularSpacing) else np.asarray(x, dtype=np.float64)


# ---- kernels with an SDE form (lti_sde.jl:176-373) ------------------------------------------------
class Kernel:
    def __add__(self, other):
        return KernelSum(_flat(KernelSum, self) + _flat(KernelSum, other))

    def __mul__(self, other):
        if isinstance(other, Kernel):
            return KernelProduct(_flat(KernelProduct, self) + _flat(KernelProduct, other))
        return ScaledKernel(self, float(other))

    def __rmul__(self, s):
        return ScaledKernel(self, float(s))


def _flat(cls, k):
    return list(k.kernels) if isinstance(k, cls) else [k]


class Matern12Kernel(Kernel):
    pass


class Matern32Kernel(Kernel):
    pass


class Matern52Kernel(Kernel):
    pass


@dataclass
class ConstantKernel(Kernel):
    c: float = 1.0


@dataclass
class ApproxPeriodicKernel(Kernel):
    N: int = 7
    r: float = 1.0


@dataclass
class ScaledKernel(Kernel):
    kernel: Kernel
    s2: float


@dataclass
class TransformedKernel(Kernel):
    """kernel ∘ ScaleTransform(s) (lti_sde.jl:346-373)."""
    kernel: Kernel
    s: float


def with_lengthscale(k: Kernel, ell: float) -> Kernel:
    return TransformedKernel(k, 1.0 / ell)


@dataclass
class KernelSum(Kernel):
    kernels: List[Kernel]


@dataclass
class KernelProduct(Kernel):
    kernels: List[Kernel]


def _sde(k):
    """(F, H, P∞) of a simple kernel: lti_sde.jl:186-235, 290-318. F, P∞ as mathematical matrices."""
    if isinstance(k, Matern12Kernel):
        return np.array([[-1.0]]), np.array([1.0]), np.array([[1.0]])
    if isinstance(k, Matern32Kernel):
        lam = math.sqrt(3.0)
        return np.array([[0.0, 1.0], [-lam * lam, -2.0 * lam]]), np.array([1.0, 0.0]), np.array([[1.0, 0.0], [0.0, 3.0]])
    if isinstance(k, Matern52Kernel):
        lam = math.sqrt(5.0)
        kap = 5.0 / 3.0
        F = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-lam ** 3, -3.0 * lam ** 2, -3.0 * lam]])
        P = np.array([[1.0, 0.0, -kap], [0.0, kap, 0.0], [-kap, 0.0, 25.0]])
        return F, np.array([1.0, 0.0, 0.0]), P
    if isinstance(k, ConstantKernel):
        return np.zeros((1, 1)), np.array([1.0]), np.array([[float(k.c)]])
    if isinstance(k, ApproxPeriodicKernel):
        N = k.N
        F = np.zeros((2 * N, 2 * N))
        P = np.zeros((2 * N, 2 * N))
        l2 = 1.0 / (4.0 * k.r ** 2)
        for i in range(N):
            w = 2.0 * math.pi * i
            F[2 * i, 2 * i + 1] = -w
            F[2 * i + 1, 2 * i] = w
            q = (1.0 if i == 0 else 2.0) * ive(i, l2)
            P[2 * i, 2 * i] = P[2 * i + 1, 2 * i + 1] = q
        return F, np.tile([1.0, 0.0], N), P
    if isinstance(k, ScaledKernel):
        F, H, P = _sde(k.kernel)
        return F, math.sqrt(k.s2) * H, P
    if isinstance(k, TransformedKernel):
        F, H, P = _sde(k.kernel)
        return F * k.s, H, P
    if isinstance(k, KernelProduct):
        F, H, P = _sde(k.kernels[0])
        for kk in k.kernels[1:]:
            F2, H2, P2 = _sde(kk)
            F = np.kron(F, np.eye(F2.shape[0])) + np.kron(np.eye(F.shape[0]), F2)
            H = np.kron(H, H2)
            P = np.kron(P, P2)
        return F, H, P
    raise TypeError(f"no state-space form for kernel {k!r}")


def _block_diag(mats):
    n = sum(m.shape[-1] for m in mats)
    lead = mats[0].shape[:-2]
    out = np.zeros(lead + (n, n))
    i = 0
    for m in mats:
        d = m.shape[-1]
        out[..., i:i + d, i:i + d] = m
        i += d
    return out


def lgssm_components(k: Kernel, t):
    """lgssm_components(k, t, storage) — lti_sde.jl:131-174 and the combinators :334-418.
    -> (As, as, Qs, Hs, hs, x0); per-step arrays are `Fill` when t is a RegularSpacing."""
    T = len(t)
    if isinstance(k, KernelSum):
        parts = [lgssm_components(kk, t) for kk in k.kernels]
        regular = all(isinstance(p[0], Fill) for p in parts)
        def cat_m(idx):
            if regular:
                return Fill(_block_diag([p[idx].value for p in parts]), T)
            return _block_diag([_dense(p[idx], T) for p in parts])
        def cat_v(idx):
            if all(isinstance(p[idx], Fill) for p in parts):
                return Fill(np.concatenate([p[idx].value for p in parts]), T)
            return np.concatenate([_dense(p[idx], T) for p in parts], axis=-1)
        hs = parts[0][4]
        for p in parts[1:]:
            hs = Fill(hs.value + p[4].value, T) if isinstance(hs, Fill) and isinstance(p[4], Fill) else _dense(hs, T) + _dense(p[4], T)
        x0 = Gaussian(np.concatenate([p[5].m for p in parts]), _block_diag([p[5].P for p in parts]))
        return cat_m(0), cat_v(1), cat_m(2), cat_v(3), hs, x0
    if isinstance(k, ScaledKernel):                        # lti_sde.jl:334-338: scales H and h
        As, as_, Qs, Hs, hs, x0 = lgssm_components(k.kernel, t)
        s = math.sqrt(k.s2)
        sc = lambda v: Fill(s * v.value, T) if isinstance(v, Fill) else s * np.asarray(v)
        return As, as_, Qs, sc(Hs), sc(hs), x0
    if isinstance(k, TransformedKernel):                   # lti_sde.jl:361-373: rescales time
        tt = RegularSpacing(k.s * t.t0, k.s * t.dt, t.N) if isinstance(t, RegularSpacing) else k.s * _times(t)
        return lgssm_components(k.kernel, tt)
    F, H, P = _sde(k)
    D = F.shape[0]
    Psym = np.triu(P) + np.triu(P, 1).T
    if isinstance(t, RegularSpacing):
        A = expm(F * t.dt)                       # ONE matrix exponential (lti_sde.jl:152)
        Q = Psym - A @ Psym @ A.T
        As, Qs = Fill(A, T), Fill(Q, T)
    else:
        tt = _times(t)
        dts = np.diff(np.concatenate([[tt[0] - 1.0], tt]))   # first transition: Δt = 1 (:139)
        uniq, inv = np.unique(dts, return_inverse=True)
        Au = np.stack([expm(F * dt) for dt in uniq])
        As = Au[inv]
        Qs = Psym - As @ Psym @ np.swapaxes(As, 1, 2)
    return As, Fill(np.zeros(D), T), Qs, Fill(H, T), Fill(np.zeros(()), T), Gaussian(np.zeros(D), P)


def sde_components(k: Kernel):
    """(F, F0, H, P) of any kernel built from the simple kernels with +, scaling and time stretching (lti_sde.jl:334-418), as ONE
    LTI SDE: block-diagonal drift F and stationary covariance P, concatenated H. F carries the time stretches (exp(F dt) is the
    transition over an UNSTRETCHED step dt); F0 is the drift without them, used by the first transition (lti_sde.jl:139, 361-373)."""
    if isinstance(k, KernelSum):
        parts = [sde_components(kk) for kk in k.kernels]
        return (_block_diag([p[0] for p in parts]), _block_diag([p[1] for p in parts]), np.concatenate([p[2] for p in parts]),
                _block_diag([p[3] for p in parts]))
    if isinstance(k, ScaledKernel):
        F, F0, H, P = sde_components(k.kernel)
        return F, F0, math.sqrt(k.s2) * H, P
    if isinstance(k, TransformedKernel):
        F, F0, H, P = sde_components(k.kernel)
        return k.s * F, F0, H, P
    F, H, P = _sde(k)
    return F, F, H, P


# Irregular grids at least this long get their transitions from the device (tgp_lti_components); shorter ones from the host loop.
DEVICE_COMPONENTS_MIN_T = 512


def lgssm_components_device(k: Kernel, t, handle):
    """lgssm_components for an irregular grid with A[t], Q[t] computed and KEPT on the GPU (tgp_lti_components): the host ships the
    time stamps (8 B/step), not the model (16 D^2 B/step), and computes no matrix exponentials."""
    import torch
    F, F0, H, P = sde_components(k)
    D, T = F.shape[0], len(t)
    dev = torch.device("cuda", handle.device)
    tt = torch.from_numpy(np.ascontiguousarray(_times(t), dtype=np.float64)).to(dev)
    A = torch.empty(T * D * D, dtype=torch.float64, device=dev)
    Q = torch.empty(T * D * D, dtype=torch.float64, device=dev)
    handle.lti_components(F, P, tt, A, Q, F0)
    return (L.DeviceSteps(A, T, (D, D)), Fill(np.zeros(D), T), L.DeviceSteps(Q, T, (D, D)), Fill(H, T), Fill(np.zeros(()), T),
            Gaussian(np.zeros(D), P))


def _dense(v, T):
    if isinstance(v, Fill):
        return np.broadcast_to(v.value, (T,) + v.value.shape).copy()
    return np.asarray(v)


# ---- separable space-time kernels on rectilinear grids (src/space_time/) -------------------------------
class SEKernel(Kernel):
    """exp(-tau^2 / 2): spatial factor only (no SDE form)."""
    def matrix(self, r):
        r = np.asarray(r, dtype=np.float64)
        return np.exp(-0.5 * (r[:, None] - r[None, :]) ** 2)


@dataclass
class Separable(Kernel):
    """separable_kernel.jl:9 — k((r, t), (r', t')) = space(r, r') * time(t, t')."""
    space: Kernel
    time: Kernel


@dataclass
class RectilinearGrid:
    """rectilinear_grid.jl:11 — xl (space) varies fastest, xr (time) slowest."""
    xl: object
    xr: object

    def __len__(self):
        return len(self.xl) * len(self.xr)


def _space_matrix(k, r):
    if hasattr(k, "matrix"):
        return k.matrix(r)
    if isinstance(k, ScaledKernel):
        return k.s2 * _space_matrix(k.kernel, r)
    if isinstance(k, TransformedKernel):
        return _space_matrix(k.kernel, k.s * np.asarray(r, dtype=np.float64))
    raise TypeError(f"no kernel matrix for spatial kernel {k!r}")


def lgssm_components_separable(k: Separable, x: RectilinearGrid):
    """lgssm_components(k::Separable, x::SpaceTimeGrid, storage) — to_gauss_markov.jl:1-24: dense Kronecker assembly,
    A = I (x) A_t, Q = (K_r + 1e-12 I) (x) Q_t, H = I (x) H_t, x0.P = K_r (x) P_t."""
    r = np.asarray(x.xl, dtype=np.float64)
    Nr = len(r)
    Kr = _space_matrix(k.space, r)
    As_t, as_t, Qs_t, Hs_t, hs_t, x0_t = lgssm_components(k.time, x.xr)
    T = len(x.xr)
    I = np.eye(Nr)

    def lift(v, f):
        return Fill(f(v.value), T) if isinstance(v, Fill) else np.stack([f(m) for m in np.asarray(v)])

    As = lift(As_t, lambda A: np.kron(I, A))
    as_ = lift(as_t, lambda a: np.tile(a, Nr))
    Qs = lift(Qs_t, lambda Q: np.kron(Kr + 1e-12 * I, Q))
    Hs = lift(Hs_t, lambda H: np.kron(I, np.atleast_2d(H)))
    hs = lift(hs_t, lambda h: np.full(Nr, float(h)))
    x0 = Gaussian(np.tile(x0_t.m, Nr), np.kron(Kr, x0_t.P))
    return As, as_, Qs, Hs, hs, x0


# ---- GP objects (AbstractGPs surface) ------------------------------------------------------------
@dataclass
class GP:
    kernel: Kernel
    mean: Union[None, float, Callable[[float], float]] = None   # ZeroMean | ConstMean | CustomMean

    def __post_init__(self):
        if isinstance(self.kernel, (int, float)) or callable(self.kernel) and not isinstance(self.kernel, Kernel):
            raise TypeError("GP(kernel[, mean])")


def _mean_vector(mean, t):
    if mean is None:
        return None
    tt = _times(t)
    if callable(mean):
        return np.array([mean(v) for v in tt], dtype=np.float64)
    return np.full(len(tt), float(mean))


@dataclass
class LTISDE:
    """lti_sde.jl:7-10."""
    f: GP
    storage: B200Storage

    def __call__(self, x, noise=1e-12):      # default noise: lti_sde.jl:27-29
        return FiniteLTISDE(self, x, noise)


def to_sde(f: GP, storage=None) -> LTISDE:
    """to_sde(f::GP, storage) — lti_sde.jl:12-14."""
    return LTISDE(f, storage or B200Storage())


def _noise_to_time_form(x, noise):
    T = len(x)
    if np.ndim(noise) == 0:
        return Fill(float(noise), T)
    noise = np.asarray(noise, dtype=np.float64)
    if noise.ndim == 2:
        noise = np.diag(noise)
    if noise.shape != (T,):
        raise L.DimensionMismatch(1, f"Dimension mismatch. length(x) is {T}, noise has shape {noise.shape}")
    return noise


def _device_components_ok(f) -> bool:
    """The device builder needs a GPU and a kernel instantiation for the composite latent dimension."""
    import torch
    if not torch.cuda.is_available():
        return False          # model construction alone stays possible on a host without a GPU; inference does not
    try:
        D = sde_components(f.f.kernel)[0].shape[0]
    except Exception:     # noqa: BLE001 — a kernel without an SDE form raises where it always did (lgssm_components)
        return False
    return D in (1, 2, 3, 4, 5, 6, 8, 10)


def build_lgssm(f: LTISDE, x, noise) -> LGSSM:
    """build_lgssm — lti_sde.jl:71-80 (+ mean handling :112-131)."""
    if isinstance(f.f.kernel, Separable):
        if f.f.mean is not None:
            raise TypeError("mean functions on space-time grids are not mirrored")
        As, as_, Qs, Hs, hs, x0 = lgssm_components_separable(f.f.kernel, x)
        Nr, T = len(x.xl), len(x.xr)
        if np.ndim(noise) == 0:     # noise_var_to_time_form(::RectilinearGrid, ::Diagonal) (rectilinear_grid.jl:90-93)
            Rs = Fill(np.full(Nr, float(noise)), T)
        else:
            Rs = np.asarray(noise, dtype=np.float64).reshape(T, Nr)
        return LGSSM(GaussMarkovModel(Forward, As, as_, Qs, x0), SmallOutputEmissions(Hs, hs, Rs))
    if not isinstance(x, RegularSpacing) and len(x) >= DEVICE_COMPONENTS_MIN_T and _device_components_ok(f):
        As, as_, Qs, Hs, hs, x0 = lgssm_components_device(f.f.kernel, x, f.storage.handle())
    else:
        As, as_, Qs, Hs, hs, x0 = lgssm_components(f.f.kernel, x)
    mean = f.f.mean
    if mean is not None and not callable(mean) and isinstance(hs, Fill):
        hs = Fill(hs.value + float(mean), len(x))          # ConstMean: mean_vector is a Fill, hs stays a Fill
    elif mean is not None:
        hs = _dense(hs, len(x)) + _mean_vector(mean, x)
    return LGSSM(GaussMarkovModel(Forward, As, as_, Qs, x0), ScalarEmissions(Hs, hs, _noise_to_time_form(x, noise)))


@dataclass
class FiniteLTISDE:
    """FiniteGP{<:LTISDE} — lti_sde.jl:24."""
    f: LTISDE
    x: object
    noise: object

    def _handle(self):
        return self.f.storage.handle()

    def build_lgssm(self):
        return build_lgssm(self.f, self.x, self.noise)


def logpdf(fx, y):
    """logpdf(ft::FiniteLTISDE, y) — lti_sde.jl:60-68; posterior variant posterior_lti_sde.jl:62-78."""
    if isinstance(fx, FinitePosteriorLTISDE):
        return _posterior_logpdf(fx, y)
    if isinstance(fx.x, RectilinearGrid):     # observations_to_time_form: space fastest (rectilinear_grid.jl:78-80)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(len(fx.x.xr), len(fx.x.xl)))
    return L.logpdf(fx.build_lgssm(), y, fx._handle())


def value_and_gradient(fun, theta, rel_step=1e-3):
    """(fun(theta), d fun / d theta) by 4th-order central differences: 4 evaluations per parameter,
    f' = (-f(+2h) + 8 f(+h) - 8 f(-h) + f(-2h)) / 12h with h = rel_step * max(1, |theta_i|)."""
    theta = np.asarray(theta, dtype=np.float64)
    f0 = float(fun(theta))
    g = np.zeros_like(theta)
    for i in range(theta.size):
        h = rel_step * max(1.0, abs(theta.flat[i]))
        def at(k):
            t = theta.copy()
            t.flat[i] += k * h
            return float(fun(t))
        g.flat[i] = (-at(2) + 8.0 * at(1) - 8.0 * at(-1) + at(-2)) / (12.0 * h)
    return f0, g


def logpdf_value_and_gradient(build, theta, y, rel_step=1e-3):
    """Gradient of the log marginal likelihood with respect to hyperparameters — what examples/exact_time_learning.jl:46-63 obtains
    by reverse-mode AD through the filter. `build(theta) -> FiniteLTISDE` (kernel, inputs and noise from the flat parameter vector).
    Here the filter is so cheap (one kernel launch over a device-resident y: 22 us per 1e7 steps) that the derivative is taken by
    4th-order central differences of the hot path itself: 4 launches per parameter, truncation O(h^4), round-off eps |lml| / h.
    y is staged on the device ONCE and shared by every evaluation. Not an adjoint: the cost grows with the number of parameters."""
    import torch
    fx0 = build(np.asarray(theta, dtype=np.float64))
    h = fx0._handle()
    if hasattr(y, "data_ptr"):
        y_dev = y
    else:
        y_dev = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).to(torch.device("cuda", h.device))
    return value_and_gradient(lambda t: logpdf(build(t), y_dev), theta, rel_step)


def marginals(fx):
    """-> (mean, var) of the marginals (lti_sde.jl:33-44, posterior_lti_sde.jl:18-37)."""
    if isinstance(fx, FinitePosteriorLTISDE):
        return _posterior_marginals(fx)
    if _is_grid(fx.x):        # marginals_diag (lti_sde.jl:39-44): the diagonal of every time step's emission covariance
        mu, v = L.marginals_diag(fx.build_lgssm(), fx._handle())
        return mu.reshape(-1), v.reshape(-1)
    return L.marginals(fx.build_lgssm(), fx._handle())


mean_and_var = marginals


def mean(fx):
    return marginals(fx)[0]


def var(fx):
    return marginals(fx)[1]


# ---- posterior (posterior_lti_sde.jl) ------------------------------------------------------------
@dataclass
class PosteriorLTISDE:
    prior: LTISDE
    y: np.ndarray
    x: object
    noise: object

    def __call__(self, x, noise=1e-12):
        return FinitePosteriorLTISDE(self, x, noise)


@dataclass
class FinitePosteriorLTISDE:
    f: PosteriorLTISDE
    x: object
    noise: object


def posterior(fx: FiniteLTISDE, y) -> PosteriorLTISDE:
    """Lazy (posterior_lti_sde.jl:7-10)."""
    if len(fx.x) != len(y):
        raise L.DimensionMismatch(1, f"Dimension mismatch. length(x) is {len(fx.x)}, but length(y) is {len(y)}")
    return PosteriorLTISDE(fx.f, y if isinstance(y, np.ma.MaskedArray) else np.asarray(y, dtype=np.float64), fx.x, fx.noise)


def _is_grid(x):
    return isinstance(x, RectilinearGrid)


def _same_inputs(a, b):
    if _is_grid(a) or _is_grid(b):
        return (_is_grid(a) and _is_grid(b) and np.array_equal(np.asarray(a.xl), np.asarray(b.xl)) and _same_inputs(a.xr, b.xr))
    if isinstance(a, RegularSpacing) and isinstance(b, RegularSpacing):
        return a == b
    ta, tb = _times(a), _times(b)
    return ta.shape == tb.shape and bool(np.all(ta == tb))


def _nan_missing(y):
    if isinstance(y, np.ma.MaskedArray):
        return y.astype(np.float64).filled(np.nan)
    return np.asarray(y, dtype=np.float64)


def _time_axis(x):
    """get_times (rectilinear_grid.jl:20)."""
    return x.xr if _is_grid(x) else x


def _obs_rows(x, y):
    """observations_to_time_form (rectilinear_grid.jl:78-80): one row per time step; space varies fastest on a grid."""
    y = _nan_missing(y)
    return y.reshape(len(x.xr), len(x.xl)) if _is_grid(x) else y


def _noise_rows(x, noise):
    """noise_var_to_time_form (lti_sde.jl:82-86, rectilinear_grid.jl:90-93) as a dense per-time-step array."""
    if not _is_grid(x):
        return _dense(_noise_to_time_form(x, noise), len(x))
    T, Nr = len(x.xr), len(x.xl)
    if np.ndim(noise) == 0:
        return np.full((T, Nr), float(noise))
    noise = np.asarray(noise, dtype=np.float64)
    if noise.ndim == 2:
        noise = np.diag(noise)
    return noise.reshape(T, Nr)


def merge_datasets(x1, x2, S1, S2, y1, y2):
    """posterior_lti_sde.jl:97-123 (+ rectilinear_grid.jl:66-75 for grids) — stable sort in time. S1, S2, y1, y2 are per-time-step
    rows (_noise_rows / _obs_rows). NaN marks missing internally; the result is handed to the LGSSM layer as a masked array."""
    x_raw = np.concatenate([_times(_time_axis(x1)), _times(_time_axis(x2))])
    idx = np.argsort(x_raw, kind="stable")
    inv = np.argsort(idx, kind="stable")
    n1, n2 = len(_time_axis(x1)), len(_time_axis(x2))
    S1, S2 = (_dense(S, n) if isinstance(S, Fill) else np.asarray(S) for S, n in ((S1, n1), (S2, n2)))
    S = np.concatenate([S1, S2])[idx]
    ys = np.concatenate([y1, y2])[idx]
    x = RectilinearGrid(x1.xl, x_raw[idx]) if _is_grid(x1) else x_raw[idx]
    return x, S, np.ma.masked_invalid(ys), inv[:n1], inv[n1:]


def _lgssm_noise(x, rows):
    """Per-step noise rows -> what build_lgssm takes for x."""
    return rows.reshape(-1) if _is_grid(x) else rows


def _posterior_marginals(fx: FinitePosteriorLTISDE):
    post = fx.f
    h = post.prior.storage.handle()
    flat = (lambda a: a.reshape(-1)) if _is_grid(fx.x) else (lambda a: a)
    if _same_inputs(fx.x, post.x):                       # posterior_lti_sde.jl:27-36
        model = build_lgssm(post.prior, post.x, post.noise)
        Rn = _noise_rows(fx.x, fx.noise) if _is_grid(fx.x) else _noise_to_time_form(fx.x, fx.noise)     # a Fill stays a Fill
        mu, v = L.posterior_marginals(model, post.y if not _is_grid(fx.x) else _masked_rows(post.x, post.y), Rn, h)
        return flat(mu), flat(v)
    n_pr = len(_time_axis(fx.x))                          # posterior_lti_sde.jl:19-26
    S_tr = _noise_rows(post.x, post.noise)
    x, S, ys, _tr, pr = merge_datasets(post.x, fx.x, S_tr, np.full((n_pr,) + S_tr.shape[1:], L.LARGE_VAR),
                                       _obs_rows(post.x, post.y), np.full((n_pr,) + S_tr.shape[1:], np.nan))
    model = build_lgssm(post.prior, x, _lgssm_noise(x, S))
    R_pr = np.zeros(S.shape)
    R_pr[pr] = _noise_rows(fx.x, fx.noise)
    mu, v = L.posterior_marginals(model, ys, R_pr, h)
    return flat(mu[pr]), flat(v[pr])


def _masked_rows(x, y):
    rows = _obs_rows(x, y)
    return np.ma.masked_invalid(rows) if np.isnan(rows).any() else rows


def _posterior_logpdf(fx: FinitePosteriorLTISDE, y_pr):
    """posterior_lti_sde.jl:62-78."""
    post = fx.f
    h = post.prior.storage.handle()
    n_pr = len(_time_axis(fx.x))
    S_pr = _noise_rows(fx.x, fx.noise)
    x, S, ys, tr, pr = merge_datasets(post.x, fx.x, _noise_rows(post.x, post.noise), S_pr, _obs_rows(post.x, post.y),
                                      np.full((n_pr,) + S_pr.shape[1:], np.nan))
    R_pr = np.zeros(S.shape)
    R_pr[pr] = S_pr
    y_full = np.full(S.shape, np.nan)
    y_full[pr] = _obs_rows(fx.x, y_pr)
    model = build_lgssm(post.prior, x, _lgssm_noise(x, S))
    model_post = L.replace_observation_noise_cov(L.posterior(model, ys, h), R_pr)
    return L.logpdf(model_post, np.ma.masked_invalid(y_full), h)
