"""Shared helpers for the parity tests: conversion between the oracle's LGSSM (test side) and the
package's host mirror (product side), random models in the style of the reference's fixtures
(test/models/model_test_utils.jl:29-58, 163-263)."""
import numpy as np

from oracle import tgp_oracle as O


def to_pkg_model(pkg, m: "O.LGSSM"):
    """oracle LGSSM -> package LGSSM (Fill where the oracle array is a stride-0 broadcast)."""
    L = pkg.lgssm
    T = m.T

    def conv(x):
        x = np.asarray(x)
        if x.strides[0] == 0 or (x.ndim and x.shape[0] == T and T > 1 and False):
            return L.Fill(np.array(x[0]), T)
        return np.array(x)

    tr = L.GaussMarkovModel(m.ordering, conv(m.As), conv(m.as_), conv(m.Qs), L.Gaussian(np.array(m.m0), np.array(m.P0)))
    em = L.ScalarEmissions(conv(m.Hs), conv(m.hs), conv(m.Rs))
    return L.LGSSM(tr, em)


def random_psd(rng, D, lo=0.1, hi=1.5):
    """PSD with eigenvalues in [lo, hi] (model_test_utils.jl:37-58)."""
    Qm, _ = np.linalg.qr(rng.standard_normal((D, D)))
    lam = rng.uniform(lo, hi, D)
    return (Qm * lam) @ Qm.T


def _stable(A, rho_max=0.97):
    """Keep the transition contractive (GP-derived LGSSMs always are); the reference's fixtures use
    N <= 49 where an eigenvalue slightly above 1 is harmless, ours run to T = 10^3..10^6."""
    rho = np.max(np.abs(np.linalg.eigvals(A)))
    return A * (rho_max / rho) if rho > rho_max else A


def random_lgssm(rng, T, D, ordering="forward", tv=True, R_lo=0.05, R_hi=1.0):
    """Random scalar-output LGSSM: A = I + 0.1 randn (model_test_utils.jl:29-31)."""
    n = T if tv else 1
    As = np.stack([_stable(0.9 * np.eye(D) + 0.1 * rng.standard_normal((D, D))) for _ in range(n)])
    as_ = rng.standard_normal((n, D)) * 0.3
    Qs = np.stack([random_psd(rng, D) for _ in range(n)])
    Hs = rng.standard_normal((n, D))
    hs = rng.standard_normal(n) * 0.2
    Rs = rng.uniform(R_lo, R_hi, n)
    if not tv:
        As, as_, Qs, Hs, hs, Rs = (np.broadcast_to(x[0], (T,) + x.shape[1:]) for x in (As, as_, Qs, Hs, hs, Rs))
    m0 = rng.standard_normal(D)
    P0 = random_psd(rng, D, 0.5, 2.0)
    return O.LGSSM(ordering, As, as_, Qs, m0, P0, Hs, hs, Rs)


def sample_y(rng, model):
    fwd = O.LGSSM("forward", model.As, model.as_, model.Qs, model.m0, model.P0, model.Hs, model.hs, model.Rs)
    return O.sample_prior(fwd, rng)


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-300))) if a.size else 0.0
