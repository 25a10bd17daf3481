"""GPU tests of the FP32-storage tensor-core variant of the large-state path (tgp_tc_gemm.cuh, tgp_dense_tc.cuh;
BASELINE config 5: ArrayStorage(Float32)). Tolerances: FP32 storage => rtol 1e-3 on logpdf against the FP64 oracle
(SURVEY.md §8d cfg 5); the contraction kernel alone must reach FP32 accuracy (3xTF32 split, ~2^-21)."""
import numpy as np
import pytest

from oracle import tgp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,Mx,N", [(32, 128, 64), (64, 128, 128), (768, 768, 768), (768, 768, 256), (256, 256, 768), (100, 70, 50),
                                    (8, 5, 3), (192, 192, 64), (33, 129, 65)])
def test_tc_gemm_matches_numpy(handle, K, Mx, N):
    """C = X' Y through TMA -> tcgen05.mma kind::tf32 (x3) -> TMEM -> tcgen05.ld, vs float64 NumPy."""
    rng = np.random.default_rng(K * 7 + Mx * 3 + N)
    X = rng.standard_normal((K, Mx)).astype(np.float32)
    Y = rng.standard_normal((K, N)).astype(np.float32)
    C = handle.tc_gemm(X, Y)
    ref = X.astype(np.float64).T @ Y.astype(np.float64)
    # The tensor core truncates every partial product to the accumulator's precision, so the error of a 3xTF32 product grows
    # linearly in K (measured on B200: 5e-7 K .. 1.2e-6 K for N(0,1) operands); a single TF32 pass would sit near 5e-4 sqrt(K).
    err = np.max(np.abs(C - ref))
    assert err < 2.5e-6 * K, f"max abs err = {err:.3e} (K = {K})"
    assert err < 1e-4 * np.sqrt(K)


@pytest.mark.parametrize("K,M", [(256, 768), (96, 200)])
def test_tc_gemm_symmetric_epilogue(handle, K, M):
    """X' X with the mirrored upper-triangle epilogue: exactly symmetric, and equal to the full product."""
    rng = np.random.default_rng(K + M)
    X = rng.standard_normal((K, M)).astype(np.float32)
    C = handle.tc_gemm(X, X, symmetric=True)
    assert np.array_equal(C, C.T)
    ref = X.astype(np.float64).T @ X.astype(np.float64)
    assert np.max(np.abs(C - ref)) < 2.5e-6 * K


def _separable(pkg, Nr, T, dtype):
    r = np.linspace(-3.0, 3.0, Nr)
    fx = pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern52Kernel())), pkg.ArrayStorage(dtype))(
        pkg.RectilinearGrid(r, pkg.RegularSpacing(0.0, 0.01, T)), 0.1)
    mo = O.build_lgssm_separable(O.SqExp(), O.Matern52(), r, O.RegularSpacing(0.0, 0.01, T), 0.1)
    return fx, mo


@pytest.mark.parametrize("Nr,T", [(24, 200), (64, 200), (256, 40)])
def test_separable_logpdf_float32_storage(pkg, Nr, T):
    """Config-5 shape (Separable(SE, Matern52), D = 3 Nr, M = Nr), FP32 storage on the tensor cores vs the FP64 oracle:
    rtol 1e-3 on logpdf, per-step values too; and against the library's own FP64 dense path."""
    fx, mo = _separable(pkg, Nr, T, np.float32)
    rng = np.random.default_rng(5 + Nr)
    y = O.sample_prior(mo, rng)
    h = fx._handle()
    try:
        lml, steps = pkg.lgssm.logpdf(fx.build_lgssm(), y, h, per_step=True)
    finally:
        h.set_dense_math(pkg.lgssm.TGP_DENSE_F64)
    ref_steps = O.logpdf_steps(mo, y)
    ref = ref_steps.sum()
    assert abs(lml - ref) <= 1e-3 * abs(ref), (lml, ref)
    np.testing.assert_allclose(steps, ref_steps, rtol=1e-3, atol=1e-3)
    # in practice the 3xTF32 path is much closer than the FP32 bar: keep it honest
    assert abs(lml - ref) <= 2e-5 * abs(ref), (lml, ref)


@pytest.mark.parametrize("ordering", ["forward", "reverse"])
def test_tc_path_time_varying_filter(pkg, handle, ordering):
    """Time-varying vector-observation model through the tensor-core step (A', H' refreshed per step), both orderings:
    filtering means / covariances and per-step lml vs the oracle at FP32 tolerance."""
    from tests.test_gpu_parity import _pkg_vector_model, _random_vector_lgssm
    rng = np.random.default_rng(77)
    T, D, M = 25, 40, 12
    m = _random_vector_lgssm(rng, T, D, M, ordering, True)
    y = O.sample_prior(O.LGSSM("forward", m.As, m.as_, m.Qs, m.m0, m.P0, m.Hs, m.hs, m.Rs), rng)
    ms_o, Ps_o, lmls_o = O.filter_(m, y)
    pm = _pkg_vector_model(pkg, m, True)
    handle.set_dense_math(pkg.lgssm.TGP_DENSE_TF32X3)
    try:
        lml, steps = pkg.lgssm.logpdf(pm, y, handle, per_step=True)
        ms, Ps = pkg.lgssm._filter(pm, y, handle)
    finally:
        handle.set_dense_math(pkg.lgssm.TGP_DENSE_F64)
    np.testing.assert_allclose(steps, lmls_o, rtol=1e-3, atol=1e-3)
    assert abs(lml - lmls_o.sum()) <= 1e-3 * abs(lmls_o.sum())
    np.testing.assert_allclose(ms, ms_o, rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(Ps, Ps_o, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-9), (np.float32, 1e-3)])
def test_dense_steady_state_switch(pkg, dtype, rtol):
    """Time-invariant large-state model, long series: after the covariance recursion has converged (device-side test) the
    library replays the mean-only step. Per-step lml must still match the sequential oracle, and the launch counter must
    show that the covariance kernels stopped running."""
    Nr, T = 16, 1800
    fx, mo = _separable(pkg, Nr, T, dtype)
    rng = np.random.default_rng(11)
    y = O.sample_prior(mo, rng)
    h = fx._handle()
    c0 = h.counters()["launches"]
    try:
        lml, steps = pkg.lgssm.logpdf(fx.build_lgssm(), y, h, per_step=True)
        c1 = h.counters()["launches"]
        h.set_algo(pkg.TGP_ALGO_SCAN)          # disables the steady-state switch: every step runs the full update
        lml_full, steps_full = pkg.lgssm.logpdf(fx.build_lgssm(), y, h, per_step=True)
        c2 = h.counters()["launches"]
    finally:
        h.set_algo(pkg.TGP_ALGO_AUTO)
        h.set_dense_math(pkg.lgssm.TGP_DENSE_F64)
    ref_steps = O.logpdf_steps(mo, y)
    atol = rtol * np.max(np.abs(ref_steps))        # lml_t crosses zero at this M: judge per-step values on the series' scale
    np.testing.assert_allclose(steps_full, ref_steps, rtol=rtol, atol=atol)
    np.testing.assert_allclose(steps, ref_steps, rtol=rtol, atol=atol)
    assert abs(lml - ref_steps.sum()) <= max(rtol, 1e-6) * abs(ref_steps.sum())
    assert (c1 - c0) < 0.8 * (c2 - c1), (c1 - c0, c2 - c1)


def test_tc_time_blocked_steady_phase(pkg):
    """FP32-storage path, log-likelihood only: after the switch the remaining steps advance block-parallel on the tensor cores
    (tc_steady_blocked: 64-step blocks, two passes + a doubling prefix over blocks). Per-step lml must agree with the
    sequential frozen replay of the same library (TGP_OPT_CHUNK = 1 disables the blocking) and with the FP64 oracle."""
    Nr, T = 20, 3000        # D = 60 (not a multiple of 32: exercises the padded stack), M = 20; ragged last block
    fx, mo = _separable(pkg, Nr, T, np.float32)
    rng = np.random.default_rng(12)
    y = O.sample_prior(mo, rng)
    h = fx._handle()
    try:
        c0 = h.counters()["launches"]
        lml_b, steps_b = pkg.lgssm.logpdf(fx.build_lgssm(), y, h, per_step=True)
        c1 = h.counters()["launches"]
        h.set_chunk(1)
        lml_s, steps_s = pkg.lgssm.logpdf(fx.build_lgssm(), y, h, per_step=True)
        c2 = h.counters()["launches"]
    finally:
        h.set_chunk(0)
        h.set_dense_math(pkg.lgssm.TGP_DENSE_F64)
    ref_steps = O.logpdf_steps(mo, y)
    scale = np.max(np.abs(ref_steps))
    np.testing.assert_allclose(steps_b, steps_s, rtol=2e-4, atol=2e-4 * scale)
    np.testing.assert_allclose(steps_b, ref_steps, rtol=1e-3, atol=1e-3 * scale)
    assert abs(lml_b - ref_steps.sum()) <= 1e-4 * abs(ref_steps.sum())
    assert (c1 - c0) < 0.5 * (c2 - c1), (c1 - c0, c2 - c1)     # ~200 big launches instead of ~6 per step
