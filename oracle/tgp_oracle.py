"""
tgp_oracle.py — CPU restatement (NumPy, FP64) of TemporalGPs.jl's LGSSM inference path.

TEST INFRASTRUCTURE ONLY. Nothing under temporalgps.jl_b200/ may import this module; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and
only as the checker.

Pinning status: the reference is pure Julia and Julia is not installed here, so the reference
cannot be executed and it ships no golden vectors. The oracle is pinned the
way the reference pins itself: every equivalence its own test-suite asserts for this path
(SDE path == dense GP for prior marginals / logpdf / posterior, test/gp/lti_sde.jl:192-201,
test/gp/posterior_lti_sde.jl:82-89, test/space_time/to_gauss_markov.jl:64-87; missing ==
analytically marginalised model, test/models/missings.jl:94-115; Scalar == 1x1 Small LGC,
test/models/linear_gaussian_conditionals.jl:117-126) is re-executed against this file in
tests/test_oracle_pins.py. All file:line citations are relative to /root/reference.

Arrays are NumPy row-major views of mathematical matrices; where the reference's column-major
layout matters (the C ABI) the conversion is done by the caller.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy.linalg import expm as _expm
from scipy.special import ive as _ive

LARGE_VAR = 1e15  # _large_var_const(), src/models/missings.jl:43
LOG2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------------
# Kernels (KernelFunctions.jl closed forms, un-vendored) and their SDE forms
# (src/gp/lti_sde.jl:176-373).
# --------------------------------------------------------------------------------------------
class Kernel:
    def __add__(self, other):
        return Sum([self, other])

    def __mul__(self, other):
        if isinstance(other, Kernel):
            return Product([self, other])
        return Scaled(float(other), self)

    def __rmul__(self, s):
        return Scaled(float(s), self)

    def stretch(self, lam):
        """k ∘ ScaleTransform(lam)"""
        return Stretched(float(lam), self)


@dataclass
class Matern12(Kernel):
    def k(self, tau):
        return np.exp(-np.abs(tau))


@dataclass
class Matern32(Kernel):
    def k(self, tau):
        r = math.sqrt(3.0) * np.abs(tau)
        return (1.0 + r) * np.exp(-r)


@dataclass
class Matern52(Kernel):
    def k(self, tau):
        r = math.sqrt(5.0) * np.abs(tau)
        return (1.0 + r + r * r / 3.0) * np.exp(-r)


@dataclass
class Constant(Kernel):
    c: float = 1.0

    def k(self, tau):
        return np.full_like(np.asarray(tau, dtype=float), self.c)


@dataclass
class ApproxPeriodic(Kernel):
    """ApproxPeriodicKernel{N}(r) (lti_sde.jl:249-320); kappa is PeriodicKernel's."""
    N: int = 7
    r: float = 1.0

    def k(self, tau):
        return np.exp(-0.5 * (np.sin(np.pi * np.asarray(tau)) / self.r) ** 2)


@dataclass
class SqExp(Kernel):
    """SEKernel exp(-tau^2/2): spatial kernels of the space-time path only (no SDE form)."""
    def k(self, tau):
        return np.exp(-0.5 * np.asarray(tau) ** 2)


@dataclass
class Scaled(Kernel):
    s2: float
    kernel: Kernel

    def k(self, tau):
        return self.s2 * self.kernel.k(tau)


@dataclass
class Stretched(Kernel):
    lam: float
    kernel: Kernel

    def k(self, tau):
        return self.kernel.k(self.lam * np.asarray(tau))


@dataclass
class Sum(Kernel):
    kernels: List[Kernel]

    def __add__(self, other):
        return Sum(self.kernels + [other])

    def k(self, tau):
        return sum(k.k(tau) for k in self.kernels)


@dataclass
class Product(Kernel):
    kernels: List[Kernel]

    def __mul__(self, other):
        if isinstance(other, Kernel):
            return Product(self.kernels + [other])
        return Scaled(float(other), self)

    def k(self, tau):
        out = 1.0
        for k in self.kernels:
            out = out * k.k(tau)
        return out


def kernelmatrix(k: Kernel, x, y=None):
    x = np.asarray(x, dtype=float)
    y = x if y is None else np.asarray(y, dtype=float)
    return k.k(x[:, None] - y[None, :])


# to_sde / stationary_distribution for base kernels. F is written as a mathematical matrix;
# the reference's SMatrix constructors are column-major (lti_sde.jl:189, 205-206, 222-223).
def to_sde(k: Kernel):
    """-> (F, q, H). lti_sde.jl:186-244, 290-304, 324-338, 346-349."""
    if isinstance(k, Matern12):
        return np.array([[-1.0]]), 2.0, np.array([1.0])
    if isinstance(k, Matern32):
        lam = math.sqrt(3.0)
        return np.array([[0.0, 1.0], [-3.0, -2.0 * lam]]), 4.0 * lam ** 3, np.array([1.0, 0.0])
    if isinstance(k, Matern52):
        lam = math.sqrt(5.0)
        F = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-lam ** 3, -3.0 * lam ** 2, -3.0 * lam]])
        return F, 8.0 * lam ** 5 / 3.0, np.array([1.0, 0.0, 0.0])
    if isinstance(k, Constant):
        return np.array([[0.0]]), 0.0, np.array([1.0])
    if isinstance(k, ApproxPeriodic):
        Fc = np.array([[0.0, -1.0], [1.0, 0.0]])  # CosineKernel F, lti_sde.jl:240
        N = k.N
        F = np.zeros((2 * N, 2 * N))
        for i in range(N):
            F[2 * i:2 * i + 2, 2 * i:2 * i + 2] = 2.0 * math.pi * i * Fc
        return F, 0.0, np.tile(np.array([1.0, 0.0]), N)
    if isinstance(k, Scaled):
        F, q, H = to_sde(k.kernel)
        s = math.sqrt(k.s2)
        return F, s * s * q, s * H
    if isinstance(k, Stretched):
        F, q, H = to_sde(k.kernel)
        return F * k.lam, q, H
    raise TypeError(f"no SDE form for {k!r}")


def stationary_distribution(k: Kernel):
    """-> (m, P). lti_sde.jl:196-201, 213-218, 230-235, 306-318, 330-332, 340-342, 351-355."""
    if isinstance(k, Matern12):
        return np.zeros(1), np.array([[1.0]])
    if isinstance(k, Matern32):
        return np.zeros(2), np.diag([1.0, 3.0])
    if isinstance(k, Matern52):
        kap = 5.0 / 3.0
        return np.zeros(3), np.array([[1.0, 0.0, -kap], [0.0, kap, 0.0], [-kap, 0.0, 25.0]])
    if isinstance(k, Constant):
        return np.zeros(1), np.array([[float(k.c)]])
    if isinstance(k, ApproxPeriodic):
        N = k.N
        l2 = 1.0 / (4.0 * k.r ** 2)
        P = np.zeros((2 * N, 2 * N))
        for j in range(1, N + 1):
            # besseli(j-1, l2)/exp(l2) == exponentially scaled Bessel ive(j-1, l2)
            qj = (1.0 + (j != 1)) * _ive(j - 1, l2)
            P[2 * (j - 1):2 * j, 2 * (j - 1):2 * j] = qj * np.eye(2)
        return np.zeros(2 * N), P
    if isinstance(k, (Scaled, Stretched)):
        return stationary_distribution(k.kernel)
    raise TypeError(f"no stationary distribution for {k!r}")


@dataclass
class RegularSpacing:
    """src/util/regular_data.jl:8-22."""
    t0: float
    dt: float
    N: int

    def collect(self):
        return self.t0 + np.arange(self.N) * self.dt

    def __len__(self):
        return self.N


def _as_times(t):
    return t.collect() if isinstance(t, RegularSpacing) else np.asarray(t, dtype=float)


def _block_diag(mats):
    n = sum(m.shape[0] for m in mats)
    k = sum(m.shape[1] for m in mats)
    out = np.zeros((n, k))
    i = j = 0
    for m in mats:
        out[i:i + m.shape[0], j:j + m.shape[1]] = m
        i += m.shape[0]
        j += m.shape[1]
    return out


def _broadcast_components(F, H, P, t):
    """broadcast_components, lti_sde.jl:131-160. Returns (As, as, Qs, Hs, hs) with leading T axis;
    the regular-spacing branch uses zero-stride broadcast views (Julia `Fill`)."""
    D = F.shape[0]
    Psym = np.triu(P) + np.triu(P, 1).T  # Symmetric(x0.P) reads the upper triangle
    if isinstance(t, RegularSpacing):
        T = t.N
        A = _expm(F * t.dt)
        Q = Psym - A @ Psym @ A.T
        As = np.broadcast_to(A, (T, D, D))
        Qs = np.broadcast_to(Q, (T, D, D))
    else:
        t = np.asarray(t, dtype=float)
        T = len(t)
        dts = np.diff(np.concatenate([[t[0] - 1.0], t]))  # first transition uses dt = 1 (:139)
        As = np.stack([_expm(F * dt) for dt in dts])
        Qs = np.stack([Psym - A @ Psym @ A.T for A in As])
    as_ = np.broadcast_to(np.zeros(D), (T, D))
    Hs = np.broadcast_to(H, (T, D))
    hs = np.broadcast_to(np.zeros(()), (T,))
    return As, as_, Qs, Hs, hs


def _apply_stretch(lam, t):
    if isinstance(t, RegularSpacing):
        return RegularSpacing(lam * t.t0, lam * t.dt, t.N)  # lti_sde.jl:373
    return lam * np.asarray(t, dtype=float)


def lgssm_components(k: Kernel, t):
    """lgssm_components(k, t, storage) for all kernel combinators, lti_sde.jl:162-174, 334-338,
    361-367, 377-397, 404-418. -> As, as, Qs, Hs, hs, (m0, P0)."""
    if isinstance(k, Scaled):
        As, as_, Qs, Hs, hs, x0 = lgssm_components(k.kernel, t)
        s = math.sqrt(k.s2)
        return As, as_, Qs, s * Hs, s * hs, x0
    if isinstance(k, Stretched):
        return lgssm_components(k.kernel, _apply_stretch(k.lam, t))
    if isinstance(k, Sum):
        parts = [lgssm_components(kk, t) for kk in k.kernels]
        T = parts[0][0].shape[0]
        As = np.stack([_block_diag([p[0][n] for p in parts]) for n in range(T)])
        as_ = np.concatenate([p[1] for p in parts], axis=1)
        Qs = np.stack([_block_diag([p[2][n] for p in parts]) for n in range(T)])
        Hs = np.concatenate([p[3] for p in parts], axis=1)
        hs = sum(p[4] for p in parts)
        m0 = np.concatenate([p[5][0] for p in parts])
        P0 = _block_diag([p[5][1] for p in parts])
        if isinstance(t, RegularSpacing):  # keep the Fill structure visible
            As = np.broadcast_to(As[0], As.shape)
            Qs = np.broadcast_to(Qs[0], Qs.shape)
        return As, as_, Qs, Hs, hs, (m0, P0)
    if isinstance(k, Product):
        sdes = [to_sde(kk) for kk in k.kernels]
        F = sdes[0][0]
        for s in sdes[1:]:
            B = s[0]
            F = np.kron(F, np.eye(B.shape[0])) + np.kron(np.eye(F.shape[0]), B)  # _kron_add :399
        H = sdes[0][2]
        for s in sdes[1:]:
            H = np.kron(H, s[2])
        x0s = [stationary_distribution(kk) for kk in k.kernels]
        m = x0s[0][0]
        P = x0s[0][1]
        for x in x0s[1:]:
            m = np.kron(m, x[0])
            P = np.kron(P, x[1])
        As, as_, Qs, Hs, hs = _broadcast_components(F, H, P, t)
        return As, as_, Qs, Hs, hs, (m, P)
    # SimpleKernel
    m0, P0 = stationary_distribution(k)
    F, _q, H = to_sde(k)  # q is never used downstream
    As, as_, Qs, Hs, hs = _broadcast_components(F, H, P0, t)
    return As, as_, Qs, Hs, hs, (m0, P0)


# --------------------------------------------------------------------------------------------
# LGSSM container (gauss_markov_model.jl:20-32, lgssm.jl:9-12)
# --------------------------------------------------------------------------------------------
@dataclass
class LGSSM:
    ordering: str            # "forward" | "reverse"
    As: np.ndarray           # (T, D, D)
    as_: np.ndarray          # (T, D)
    Qs: np.ndarray           # (T, D, D)
    m0: np.ndarray           # (D,)
    P0: np.ndarray           # (D, D)
    Hs: np.ndarray           # scalar emissions: (T, D); vector emissions: (T, M, D)
    hs: np.ndarray           # (T,) | (T, M)
    Rs: np.ndarray           # (T,) | (T, M, M)

    @property
    def T(self):
        return self.As.shape[0]

    @property
    def D(self):
        return self.As.shape[1]

    @property
    def scalar(self):
        return self.Hs.ndim == 2

    def indices(self):
        """eachindex: 1:T forward, T:-1:1 reverse (gauss_markov_model.jl:36-40); 0-based here."""
        return range(self.T) if self.ordering == "forward" else range(self.T - 1, -1, -1)


def build_lgssm(k: Kernel, t, sigma2, mean=None) -> LGSSM:
    """build_lgssm, lti_sde.jl:71-80 (+ mean handling :119-131). `sigma2` scalar or (T,) vector;
    `mean` None (ZeroMean), a float (ConstMean) or a callable (CustomMean)."""
    As, as_, Qs, Hs, hs, (m0, P0) = lgssm_components(k, t)
    tt = _as_times(t)
    T = len(tt)
    if mean is not None:
        mv = np.full(T, float(mean)) if not callable(mean) else np.array([mean(x) for x in tt])
        hs = hs + mv
    Rs = np.broadcast_to(np.asarray(sigma2, dtype=float), (T,))
    return LGSSM("forward", As, as_, Qs, m0, P0, Hs, hs, Rs)


# --------------------------------------------------------------------------------------------
# One-step math (src/models/linear_gaussian_conditionals.jl)
# --------------------------------------------------------------------------------------------
def _symmetric(P):
    return np.triu(P) + np.triu(P, 1).T


def predict(m, P, A, a, Q):
    """LGC:46-52 — A m + a, (A * Symmetric(P)) * A' + Q (upper triangle of P only)."""
    return A @ m + a, (A @ _symmetric(P)) @ A.T + Q


def posterior_and_lml_scalar(m, P, H, h, R, y):
    """ScalarOutputLGC, LGC:247-257."""
    V = H @ P
    sqrtS = math.sqrt(V @ H + R)
    B = V / sqrtS
    alpha = (y - (H @ m + h)) / sqrtS
    lml = -(LOG2PI + 2.0 * math.log(sqrtS) + alpha * alpha) / 2.0
    return m + B * alpha, P - np.outer(B, B), lml


def posterior_and_lml_small(m, P, H, h, R, y):
    """SmallOutputLGC, LGC:129-141 (Cholesky of Symmetric(V A' + Q), upper triangle)."""
    V = H @ P
    S = _symmetric(V @ H.T + R)
    U = np.linalg.cholesky(S).T  # S = U'U
    B = np.linalg.solve(U.T, V)
    alpha = np.linalg.solve(U.T, y - (H @ m + h))
    logdet = 2.0 * np.sum(np.log(np.diag(U)))
    lml = -(len(y) * LOG2PI + logdet + alpha @ alpha) / 2.0
    return m + B.T @ alpha, P - B.T @ B, lml


def posterior_and_lml_large(m, P, H, h, R, y):
    """LargeOutputLGC, LGC:179-204 (jitter 1e-10 on P)."""
    D = len(m)
    UQ = np.linalg.cholesky(_symmetric(R)).T
    UP = np.linalg.cholesky(_symmetric(P + 1e-10 * np.eye(D))).T
    Bt = np.linalg.solve(UQ.T, H) @ UP.T
    UF = np.linalg.cholesky(_symmetric(Bt.T @ Bt + np.eye(D))).T
    G = np.linalg.solve(UF.T, UP)
    P_post = G.T @ G
    delta = np.linalg.solve(UQ.T, y - (H @ m + h))
    beta = np.linalg.solve(UF.T, Bt.T @ delta)
    m_post = m + G.T @ beta
    c = len(y) * LOG2PI
    logdetF = 2.0 * np.sum(np.log(np.diag(UF)))
    logdetQ = 2.0 * np.sum(np.log(np.diag(UQ)))
    lml = -(delta @ delta - beta @ beta + c + logdetF + logdetQ) / 2.0
    return m_post, P_post, lml


def posterior_and_lml_bottleneck(m, P, H, h, A, a, Q, y):
    """BottleneckLGC(H, h, LargeOutputLGC(A, a, Q)), LGC:305-335: project onto z = H x + h (jitter 1e-12, _project :305-309),
    condition z with the fan-out LargeOutputLGC, then integrate x | z against z | y (jitter 1e-12 in the Cholesky, :330)."""
    K = H.shape[0]
    zm, zP = H @ m + h, H @ P @ H.T + 1e-12 * np.eye(K)
    zpm, zpP, lml = posterior_and_lml_large(zm, zP, A, a, Q, y)
    U = np.linalg.cholesky(_symmetric(zP + 1e-12 * np.eye(K))).T
    Gt = np.linalg.solve(U, np.linalg.solve(U.T, H @ P))
    return m + Gt.T @ (zpm - zm), P + Gt.T @ (zpP - zP) @ Gt, lml


def _update(model: LGSSM, t, m, P, y):
    if model.scalar:
        return posterior_and_lml_scalar(m, P, model.Hs[t], float(model.hs[t]), float(model.Rs[t]), float(y))
    return posterior_and_lml_small(m, P, model.Hs[t], model.hs[t], model.Rs[t], np.asarray(y))


def _emit_predict(model: LGSSM, t, m, P):
    """predict(x, emission) (LGC:46-52) in emission space."""
    if model.scalar:
        H = model.Hs[t]
        return H @ m + model.hs[t], (H @ _symmetric(P)) @ H + model.Rs[t]
    H = model.Hs[t]
    return H @ m + model.hs[t], (H @ _symmetric(P)) @ H.T + model.Rs[t]


# --------------------------------------------------------------------------------------------
# Recursions (src/models/lgssm.jl)
# --------------------------------------------------------------------------------------------
def pairwise_sum(x):
    """Julia's sum(::Vector{Float64}) is pairwise with 1024-element leaves (Base.mapreduce_impl)."""
    x = np.asarray(x, dtype=float)
    n = len(x)
    if n <= 1024:
        s = 0.0
        for v in x:
            s += float(v)
        return s
    h = n // 2
    return pairwise_sum(x[:h]) + pairwise_sum(x[h:])


def logpdf_steps(model: LGSSM, y):
    """scan_emit(step_logpdf, ...) lgssm.jl:147-165; returns the per-step lml vector (memory order)."""
    m, P = model.m0, model.P0
    lmls = np.empty(model.T)
    for t in model.indices():
        if model.ordering == "forward":
            m, P = predict(m, P, model.As[t], model.as_[t], model.Qs[t])
            m, P, lml = _update(model, t, m, P, y[t])
        else:
            m, P, lml = _update(model, t, m, P, y[t])
            m, P = predict(m, P, model.As[t], model.as_[t], model.Qs[t])
        lmls[t] = lml
    return lmls


def logpdf(model: LGSSM, y):
    return pairwise_sum(logpdf_steps(model, y))


def filter_(model: LGSSM, y):
    """_filter, lgssm.jl:171-187 -> (ms (T,D), Ps (T,D,D), lmls (T,)) in memory order."""
    m, P = model.m0, model.P0
    ms = np.empty((model.T, model.D))
    Ps = np.empty((model.T, model.D, model.D))
    lmls = np.empty(model.T)
    for t in model.indices():
        if model.ordering == "forward":
            m, P = predict(m, P, model.As[t], model.as_[t], model.Qs[t])
            m, P, lmls[t] = _update(model, t, m, P, y[t])
            ms[t], Ps[t] = m, P
        else:
            m, P, lmls[t] = _update(model, t, m, P, y[t])
            ms[t], Ps[t] = m, P
            m, P = predict(m, P, model.As[t], model.as_[t], model.Qs[t])
    return ms, Ps, lmls


def invert_dynamics(mf, Pf, mp, Pp, A):
    """lgssm.jl:231-240: U = chol(Symmetric(Pp + 1e-10 I)); Gt = U \\ (U' \\ (A Pf));
    returns (G, g, Sigma) with Sigma = Pf - (U Gt)'(U Gt)."""
    D = len(mf)
    U = np.linalg.cholesky(_symmetric(Pp + 1e-10 * np.eye(D))).T
    Gt = np.linalg.solve(U, np.linalg.solve(U.T, A @ Pf))
    B = U @ Gt
    return Gt.T, mf - Gt.T @ mp, Pf - B.T @ B


def posterior(model: LGSSM, y) -> LGSSM:
    """posterior(::LGSSM, y), lgssm.jl:193-228. Returns the reversed-ordering LGSSM."""
    if model.T != len(y):
        raise ValueError(f"Dimension mismatch. length(prior) is {model.T}, but length(y) is {len(y)}")
    T, D = model.T, model.D
    G = np.empty((T, D, D))
    g = np.empty((T, D))
    S = np.empty((T, D, D))
    m, P = model.m0, model.P0
    for t in model.indices():
        A, a, Q = model.As[t], model.as_[t], model.Qs[t]
        if model.ordering == "forward":
            mp, Pp = predict(m, P, A, a, Q)
            G[t], g[t], S[t] = invert_dynamics(m, P, mp, Pp, A)
            m, P, _ = _update(model, t, mp, Pp, y[t])
        else:
            mf, Pf, _ = _update(model, t, m, P, y[t])
            mp, Pp = predict(mf, Pf, A, a, Q)
            # step_posterior(::Reverse) calls invert_dynamics(xp, xf, t) (lgssm.jl:227): swapped.
            G[t], g[t], S[t] = invert_dynamics(mp, Pp, mf, Pf, A)
            m, P = mp, Pp
    new_order = "reverse" if model.ordering == "forward" else "forward"
    return LGSSM(new_order, G, g, S, m, P, model.Hs, model.hs, model.Rs)


def marginals(model: LGSSM):
    """marginals(::LGSSM), lgssm.jl:99-115 -> emission-space (means, covs) in memory order."""
    m, P = model.m0, model.P0
    T = model.T
    if model.scalar:
        means = np.empty(T)
        covs = np.empty(T)
    else:
        M = model.Hs.shape[1]
        means = np.empty((T, M))
        covs = np.empty((T, M, M))
    for t in model.indices():
        if model.ordering == "forward":
            m, P = predict(m, P, model.As[t], model.as_[t], model.Qs[t])
            means[t], covs[t] = _emit_predict(model, t, m, P)
        else:
            means[t], covs[t] = _emit_predict(model, t, m, P)
            m, P = predict(m, P, model.As[t], model.as_[t], model.Qs[t])
    return means, covs


def marginals_diag(model: LGSSM):
    """marginals_diag(::LGSSM), lgssm.jl:125-141 with predict_marginals (linear_gaussian_conditionals.jl:63-68): the same recursion
    as marginals(), emitting Gaussian(H m + h, Diagonal(diag(H P H') + diag(R)))."""
    means, covs = marginals(model)
    if model.scalar:
        return means, covs
    return means, np.einsum("tii->ti", covs).copy()


def replace_observation_noise_cov(model: LGSSM, Rs_new) -> LGSSM:
    """missings.jl:35-41."""
    Rs_new = np.asarray(Rs_new, dtype=float)
    if model.scalar:
        Rs_new = np.broadcast_to(Rs_new, (model.T,))
    return LGSSM(model.ordering, model.As, model.as_, model.Qs, model.m0, model.P0, model.Hs, model.hs, Rs_new)


# --------------------------------------------------------------------------------------------
# Missing data (src/models/missings.jl). Missing observations are NaN in `y`.
# --------------------------------------------------------------------------------------------
def transform_model_and_obs(model: LGSSM, y):
    """missings.jl:25-33, 59-74: y := 0, R := 1e15 at missing steps (whole-observation missing)."""
    y = np.array(y, dtype=float, copy=True)
    if model.scalar:
        miss = np.isnan(y)
        Rs = np.array(np.broadcast_to(model.Rs, (model.T,)), copy=True)
        Rs[miss] = LARGE_VAR
        y[miss] = 0.0
        n_missing_dims = int(miss.sum())
    else:
        nan = np.isnan(y)
        miss = nan.all(axis=1)
        M = y.shape[1]
        Rs = np.array(np.broadcast_to(model.Rs, (model.T, M, M)), copy=True)
        Rs[miss] = LARGE_VAR * np.eye(M)
        n_missing_dims = int(miss.sum()) * M
        # element-wise missing (linear_gaussian_conditionals.jl:143-151 + missings.jl:76-79): Diagonal R only; the
        # missing entries get variance 1e15 and observation 0, each adds log(2 pi 1e15) / 2
        for t in np.nonzero(nan.any(axis=1) & ~miss)[0]:
            if np.abs(Rs[t] - np.diag(np.diag(Rs[t]))).max() != 0.0:
                raise TypeError("MethodError: element-wise missing observations need a Diagonal observation covariance")
            for i in np.nonzero(nan[t])[0]:
                Rs[t, i, i] = LARGE_VAR
                n_missing_dims += 1
        y[nan] = 0.0
    return replace_observation_noise_cov(model, Rs), y, n_missing_dims


def logpdf_missing(model: LGSSM, y):
    """missings.jl:8-13, 45-49."""
    model2, y2, n = transform_model_and_obs(model, y)
    return logpdf(model2, y2) + n * math.log(2.0 * math.pi * LARGE_VAR) / 2.0


# --------------------------------------------------------------------------------------------
# GP-level API (src/gp/lti_sde.jl:33-68, src/gp/posterior_lti_sde.jl)
# --------------------------------------------------------------------------------------------
def gp_logpdf(k, t, sigma2, y, mean=None):
    """logpdf(f(t, sigma2), y) via the SDE path; NaN in y = missing."""
    model = build_lgssm(k, t, sigma2, mean)
    y = np.asarray(y, dtype=float)
    if np.isnan(y).any():
        return logpdf_missing(model, y)
    return logpdf(model, y)


def gp_prior_marginals(k, t, sigma2, mean=None):
    """marginals(ft::FiniteLTISDE), lti_sde.jl:33-44 -> (mean, var)."""
    return marginals(build_lgssm(k, t, sigma2, mean))


def merge_datasets(x1, x2, S1, S2, y1, y2):
    """posterior_lti_sde.jl:97-123 (NaN = missing). Stable merge by sortperm."""
    x_raw = np.concatenate([x1, x2])
    idx = np.argsort(x_raw, kind="stable")
    x = x_raw[idx]
    Sig = np.concatenate([S1, S2])[idx]
    ys = np.concatenate([y1, y2])[idx]
    inv = np.argsort(idx, kind="stable")
    return x, Sig, ys, inv[:len(x1)], inv[len(x1):]


def gp_posterior_marginals(k, t, sigma2, y, t_pr=None, sigma2_pr=1e-12, mean=None):
    """marginals(f_post(t_pr, sigma2_pr)), posterior_lti_sde.jl:18-37 -> (mean, var).
    t_pr None => same inputs branch (:27-36)."""
    tt = _as_times(t)
    y = np.asarray(y, dtype=float)
    if t_pr is None:
        model = build_lgssm(k, t, sigma2, mean)
        post = posterior_missing(model, y)
        return marginals(replace_observation_noise_cov(post, np.broadcast_to(sigma2_pr, (len(tt),))))
    t_pr = np.asarray(t_pr, dtype=float)
    S1 = np.broadcast_to(np.asarray(sigma2, dtype=float), (len(tt),))
    x, Sig, ys, _tr, pr = merge_datasets(tt, t_pr, S1, np.full(len(t_pr), LARGE_VAR), y,
                                         np.full(len(t_pr), np.nan))
    model = build_lgssm(k, x, Sig, mean)
    Rs_pr_full = np.zeros(len(x))
    Rs_pr_full[pr] = np.broadcast_to(sigma2_pr, (len(t_pr),))
    post = replace_observation_noise_cov(posterior_missing(model, ys), Rs_pr_full)
    mu, var = marginals(post)
    return mu[pr], var[pr]


def posterior_missing(model, y):
    """missings.jl:20-23."""
    if np.isnan(y).any():
        model2, y2, _ = transform_model_and_obs(model, y)
        return posterior(model2, y2)
    return posterior(model, y)


def gp_posterior_logpdf(k, t, sigma2, y, t_pr, sigma2_pr, y_pr, mean=None):
    """logpdf(f_post(t_pr, sigma2_pr), y_pr), posterior_lti_sde.jl:62-78."""
    tt = _as_times(t)
    t_pr = np.asarray(t_pr, dtype=float)
    S1 = np.broadcast_to(np.asarray(sigma2, dtype=float), (len(tt),))
    S2 = np.broadcast_to(np.asarray(sigma2_pr, dtype=float), (len(t_pr),))
    x, Sig, ys, tr, pr = merge_datasets(tt, t_pr, S1, S2, np.asarray(y, dtype=float),
                                        np.full(len(t_pr), np.nan))
    Rs_pr_full = np.zeros(len(x))
    Rs_pr_full[pr] = S2
    ys_pr_full = np.full(len(x), np.nan)
    ys_pr_full[pr] = y_pr
    model = build_lgssm(k, x, Sig, mean)
    post = replace_observation_noise_cov(posterior_missing(model, ys), Rs_pr_full)
    return logpdf_missing(post, ys_pr_full)


# --------------------------------------------------------------------------------------------
# Dense ("naive") GP — the AbstractGPs side of the reference's equivalence tests.
# --------------------------------------------------------------------------------------------
def _mean_vec(mean, x):
    x = np.asarray(x, dtype=float)
    if mean is None:
        return np.zeros(len(x))
    if callable(mean):
        return np.array([mean(v) for v in x])
    return np.full(len(x), float(mean))


def dense_logpdf(k, t, sigma2, y, mean=None):
    tt = _as_times(t)
    K = kernelmatrix(k, tt) + np.diag(np.broadcast_to(sigma2, (len(tt),)))
    L = np.linalg.cholesky(K)
    r = np.linalg.solve(L, np.asarray(y) - _mean_vec(mean, tt))
    return -0.5 * (len(tt) * LOG2PI + 2.0 * np.sum(np.log(np.diag(L))) + r @ r)


def dense_prior_marginals(k, t, sigma2, mean=None):
    tt = _as_times(t)
    return _mean_vec(mean, tt), np.diag(kernelmatrix(k, tt)) + np.broadcast_to(sigma2, (len(tt),))


def dense_posterior(k, t, sigma2, y, t_pr, sigma2_pr, mean=None):
    """-> (mean, cov) of f_post(t_pr, sigma2_pr) incl. predictive noise on the diagonal."""
    tt = _as_times(t)
    t_pr = np.asarray(t_pr, dtype=float)
    K = kernelmatrix(k, tt) + np.diag(np.broadcast_to(sigma2, (len(tt),)))
    Ks = kernelmatrix(k, t_pr, tt)
    Kss = kernelmatrix(k, t_pr)
    L = np.linalg.cholesky(K)
    alpha = np.linalg.solve(L.T, np.linalg.solve(L, np.asarray(y) - _mean_vec(mean, tt)))
    V = np.linalg.solve(L, Ks.T)
    mu = _mean_vec(mean, t_pr) + Ks @ alpha
    cov = Kss - V.T @ V + np.diag(np.broadcast_to(sigma2_pr, (len(t_pr),)))
    return mu, cov


def dense_posterior_logpdf(k, t, sigma2, y, t_pr, sigma2_pr, y_pr, mean=None):
    mu, cov = dense_posterior(k, t, sigma2, y, t_pr, sigma2_pr, mean)
    L = np.linalg.cholesky(cov)
    r = np.linalg.solve(L, np.asarray(y_pr) - mu)
    return -0.5 * (len(mu) * LOG2PI + 2.0 * np.sum(np.log(np.diag(L))) + r @ r)


# --------------------------------------------------------------------------------------------
# Space-time separable models (src/space_time/to_gauss_markov.jl:1-24)
# --------------------------------------------------------------------------------------------
def build_lgssm_separable(k_space: Kernel, k_time: Kernel, r, t, sigma2) -> LGSSM:
    """Separable(k_space, k_time) on RectilinearGrid(r, t): dense kron assembly (:13-18);
    observations per time step are the Nr spatial points (space fastest, rectilinear_grid.jl:33-35)."""
    r = np.asarray(r, dtype=float)
    Nr = len(r)
    Kr = kernelmatrix(k_space, r)
    As_t, as_t, Qs_t, Hs_t, hs_t, (m0_t, P0_t) = lgssm_components(k_time, t)
    T = As_t.shape[0]
    I = np.eye(Nr)
    regular = isinstance(t, RegularSpacing)

    def lift(mats, f):
        if regular:
            v = f(mats[0])
            return np.broadcast_to(v, (T,) + v.shape)
        return np.stack([f(m) for m in mats])

    As = lift(As_t, lambda A: np.kron(I, A))
    as_ = lift(as_t, lambda a: np.tile(a, Nr))
    Qs = lift(Qs_t, lambda Q: np.kron(Kr + 1e-12 * I, Q))
    Hs = lift(Hs_t, lambda H: np.kron(I, H[None, :]))
    hs = np.stack([np.full(Nr, float(h)) for h in hs_t]) if not regular else np.broadcast_to(
        np.full(Nr, float(hs_t[0])), (T, Nr))
    m0 = np.tile(m0_t, Nr)
    P0 = np.kron(Kr, P0_t)
    s2 = np.broadcast_to(np.asarray(sigma2, dtype=float), (T * Nr,)).reshape(T, Nr)
    Rs = np.stack([np.diag(v) for v in s2]) if not regular or np.ndim(sigma2) else np.broadcast_to(
        np.diag(s2[0]), (T, Nr, Nr))
    return LGSSM("forward", As, as_, Qs, m0, P0, Hs, hs, Rs)


def dense_separable_posterior(k_space, k_time, r, t, sigma2, y, t_pr, sigma2_pr):
    """Dense-GP posterior of a Separable(k_space, k_time) GP observed on RectilinearGrid(r, t) at RectilinearGrid(r, t_pr):
    -> (mean, covariance incl. sigma2_pr), points ordered space-fastest (the naive side of test/space_time/to_gauss_markov.jl:68-87)."""
    r = np.asarray(r, dtype=float)
    tt, tp = _as_times(t), _as_times(t_pr)
    Kr = kernelmatrix(k_space, r)
    K = np.kron(kernelmatrix(k_time, tt), Kr) + sigma2 * np.eye(len(tt) * len(r))
    Ks = np.kron(kernelmatrix(k_time, tp, tt), Kr)
    Kss = np.kron(kernelmatrix(k_time, tp), Kr)
    Lc = np.linalg.cholesky(K)
    alpha = np.linalg.solve(Lc.T, np.linalg.solve(Lc, np.asarray(y, dtype=float).reshape(-1)))
    V = np.linalg.solve(Lc, Ks.T)
    return Ks @ alpha, Kss - V.T @ V + sigma2_pr * np.eye(len(tp) * len(r))


def dense_separable_logpdf(k_space, k_time, r, t, sigma2, y):
    """y flat with space fastest."""
    r = np.asarray(r, dtype=float)
    tt = _as_times(t)
    K = np.kron(kernelmatrix(k_time, tt), kernelmatrix(k_space, r))
    K = K + np.diag(np.broadcast_to(sigma2, (len(tt) * len(r),)))
    L = np.linalg.cholesky(K)
    v = np.linalg.solve(L, np.asarray(y))
    return -0.5 * (len(y) * LOG2PI + 2.0 * np.sum(np.log(np.diag(L))) + v @ v)


# --------------------------------------------------------------------------------------------
# Sampling from the model itself (synthetic inputs; not RNG-compatible with Julia)
# --------------------------------------------------------------------------------------------
def sample_prior(model: LGSSM, rng: np.random.Generator):
    """Ancestral sample of y from a Forward LGSSM (structure of lgssm.jl:65-85; own RNG stream)."""
    assert model.ordering == "forward"
    D = model.D
    x = model.m0 + np.linalg.cholesky(_symmetric(model.P0) + 1e-12 * np.eye(D)) @ rng.standard_normal(D)
    ys = []
    cache = {}
    for t in range(model.T):
        Q = model.Qs[t]
        key = id(Q.base) if Q.base is not None and model.Qs.strides[0] == 0 else None
        if key is not None and key in cache:
            LQ = cache[key]
        else:
            LQ = np.linalg.cholesky(_symmetric(Q) + 1e-9 * np.eye(D))
            if key is not None:
                cache[key] = LQ
        x = model.As[t] @ x + model.as_[t] + LQ @ rng.standard_normal(D)
        if model.scalar:
            ys.append(model.Hs[t] @ x + model.hs[t] + math.sqrt(model.Rs[t]) * rng.standard_normal())
        else:
            M = model.Hs.shape[1]
            LR = np.linalg.cholesky(_symmetric(model.Rs[t]))
            ys.append(model.Hs[t] @ x + model.hs[t] + LR @ rng.standard_normal(M))
    return np.array(ys)
