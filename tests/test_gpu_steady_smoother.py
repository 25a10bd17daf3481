"""GPU tests of the steady-state posterior-marginals path (tgp_steady_smooth.cuh): time-invariant models, long series — head and
tail by the general scan kernels, the rest by constant-coefficient scans over vectors. Compared with the sequential C oracle
(filter + RTS smoother, oracle/lgssm_ref.c) at north_star's tolerances: 1e-6 relative on logpdf, 1e-5 on means / variances."""
import numpy as np
import pytest

from oracle import c_oracle, tgp_oracle as O

pytestmark = pytest.mark.gpu
LML_RTOL, MV_RTOL = 1e-6, 1e-5


def _cfg3_kernels(pkg):
    TK = pkg.gp.TransformedKernel
    kp = 1.0 * pkg.Matern32Kernel() + 0.7 * pkg.Matern52Kernel() + 0.5 * TK(pkg.Matern52Kernel(), 0.5) + 0.3 * TK(pkg.Matern32Kernel(), 2.0)
    ko = 1.0 * O.Matern32() + 0.7 * O.Matern52() + 0.5 * O.Matern52().stretch(0.5) + 0.3 * O.Matern32().stretch(2.0)
    return kp, ko


def _run(pkg, handle, kp, ko, T, dt, sigma2, r_new, seed):
    mo = O.build_lgssm(ko, O.RegularSpacing(0.0, dt, T), sigma2)
    rng = np.random.default_rng(seed)
    y = np.sin(np.arange(T) * 0.004) + 0.3 * np.cos(np.arange(T) * 0.05) + 0.35 * rng.standard_normal(T)
    model = pkg.to_sde(pkg.GP(kp))(pkg.RegularSpacing(0.0, dt, T), sigma2).build_lgssm()
    c0 = handle.counters()["launches"]
    mu, var, lml = pkg.lgssm.posterior_marginals(model, y, r_new, handle, return_lml=True)
    launches = handle.counters()["launches"] - c0
    mu_o, var_o, lml_o = c_oracle.posterior_marginals(c_oracle.Model.from_lgssm(mo), y, r_new)
    return (mu, var, lml), (mu_o, var_o, lml_o), launches, (model, y)


@pytest.mark.parametrize("name,T", [("matern52", 40_000), ("matern32", 33_001), ("cfg3", 50_000), ("matern12", 70_007)])
def test_steady_posterior_marginals_match_oracle(pkg, handle, name, T):
    if name == "cfg3":
        kp, ko = _cfg3_kernels(pkg)           # D = 10 (BASELINE config 3's stand-in kernel)
    else:
        kp = {"matern52": pkg.Matern52Kernel, "matern32": pkg.Matern32Kernel, "matern12": pkg.Matern12Kernel}[name]()
        ko = {"matern52": O.Matern52, "matern32": O.Matern32, "matern12": O.Matern12}[name]()
    got, ref, launches, (model, y) = _run(pkg, handle, kp, ko, T, 0.01, 0.1, 1e-2, 31)
    np.testing.assert_allclose(got[0], ref[0], rtol=MV_RTOL, atol=1e-7)
    np.testing.assert_allclose(got[1], ref[1], rtol=MV_RTOL)
    assert abs(got[2] - ref[2]) <= LML_RTOL * abs(ref[2])
    # the steady path was taken (it launches the constant-coefficient scan kernels: more launches than the general path's ~12)
    handle.set_algo(pkg.TGP_ALGO_SCAN)
    try:
        c0 = handle.counters()["launches"]
        mu_g, var_g, lml_g = pkg.lgssm.posterior_marginals(model, y, 1e-2, handle, return_lml=True)
        launches_general = handle.counters()["launches"] - c0
    finally:
        handle.set_algo(pkg.TGP_ALGO_AUTO)
    assert launches > launches_general
    np.testing.assert_allclose(got[0], mu_g, rtol=MV_RTOL, atol=1e-7)
    np.testing.assert_allclose(got[1], var_g, rtol=MV_RTOL)


def test_steady_posterior_marginals_heteroscedastic_prediction_noise(pkg, handle):
    """R_new per step (stride 1): only the emitted variances change."""
    T = 36_000
    rng = np.random.default_rng(5)
    r_new = rng.uniform(0.01, 0.5, T)
    got, ref, _, _ = _run(pkg, handle, pkg.Matern52Kernel(), O.Matern52(), T, 0.01, 0.1, r_new, 32)
    np.testing.assert_allclose(got[0], ref[0], rtol=MV_RTOL, atol=1e-7)
    np.testing.assert_allclose(got[1], ref[1], rtol=MV_RTOL)


def test_steady_posterior_marginals_falls_back_when_not_converged(pkg, handle):
    """A very fine grid: the covariances have not converged inside the 4096-step head / tail, the device-side test fails and the
    call is redone by the general kernels — same answer."""
    T = 40_000
    got, ref, launches, _ = _run(pkg, handle, pkg.Matern52Kernel(), O.Matern52(), T, 1e-5, 0.1, 1e-2, 33)
    np.testing.assert_allclose(got[0], ref[0], rtol=MV_RTOL, atol=1e-7)
    np.testing.assert_allclose(got[1], ref[1], rtol=MV_RTOL)
    assert abs(got[2] - ref[2]) <= LML_RTOL * abs(ref[2])


def test_steady_logpdf_large_state_dimension(pkg, handle):
    """logpdf of the D = 10 config-3 kernel on a long regular grid: no register-resident steady kernel exists for D > 6, so the
    library runs the sequential single-CTA head + ONE constant-coefficient forward scan (logpdf_steady_vec). Same answer as the
    general scan (TGP_ALGO_SCAN) and the sequential oracle, with fewer, lighter launches."""
    kp, ko = _cfg3_kernels(pkg)
    T = 60_000
    mo = O.build_lgssm(ko, O.RegularSpacing(0.0, 0.01, T), 0.1)
    rng = np.random.default_rng(77)
    y = np.sin(np.arange(T) * 0.004) + 0.35 * rng.standard_normal(T)
    model = pkg.to_sde(pkg.GP(kp))(pkg.RegularSpacing(0.0, 0.01, T), 0.1).build_lgssm()
    ref = c_oracle.logpdf(c_oracle.Model.from_lgssm(mo), y)
    lml = pkg.lgssm.logpdf(model, y, handle)
    assert abs(lml - ref) <= LML_RTOL * abs(ref)
    handle.set_algo(pkg.TGP_ALGO_SCAN)
    try:
        lml_g = pkg.lgssm.logpdf(model, y, handle)
    finally:
        handle.set_algo(pkg.TGP_ALGO_AUTO)
    assert abs(lml - lml_g) <= 1e-9 * abs(ref)
    # per-step output requested: the general kernels (the vector scan emits the sum only)
    lml2, steps = pkg.lgssm.logpdf(model, y, handle, per_step=True)
    assert abs(lml2 - ref) <= LML_RTOL * abs(ref) and abs(steps.sum() - ref) <= LML_RTOL * abs(ref)


@pytest.mark.parametrize("world", [2, 5])
def test_time_sharded_posterior_marginals_need_no_exchange(pkg, world):
    """sharded.posterior_marginals_sharded: every rank smooths its shard extended by a halo on both sides; the concatenation of the
    interiors equals the single-call posterior marginals of the whole series and the sequential oracle (1e-5) — no communication."""
    from temporalgps_jl_b200 import sharded
    T, dt, s2 = 30_000, 0.01, 0.1
    rng = np.random.default_rng(23)
    x = pkg.RegularSpacing(0.0, dt, T)
    y = np.sin(np.arange(T) * 0.004) + 0.35 * rng.standard_normal(T)
    f = pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))
    mu_all, var_all = pkg.gp.marginals(pkg.gp.posterior(f(x, s2), y)(x, 1e-2))
    pieces = []
    for r in range(world):
        lo, hi, lo_h, hi_h = sharded.halo_bounds(T, world, r, 4096)
        pieces.append(sharded.posterior_marginals_sharded(pkg.gp, f, x, s2, y[lo_h:hi_h], 1e-2, r, world))
        assert len(pieces[-1][0]) == hi - lo
    mu = np.concatenate([p[0] for p in pieces])
    var = np.concatenate([p[1] for p in pieces])
    np.testing.assert_allclose(mu, mu_all, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var, var_all, rtol=1e-9)
    mu_o, var_o = O.gp_posterior_marginals(O.Matern52(), O.RegularSpacing(0.0, dt, T), s2, y, None, 1e-2)
    np.testing.assert_allclose(mu, mu_o, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(var, var_o, rtol=1e-5)
