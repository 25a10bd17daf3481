"""GPU parity tests: the CUDA path (through the C ABI of libtgpb200.so) against the CPU oracle on the
same seeded inputs. Tolerances are north_star's: 1e-6 relative on logpdf, 1e-5 on means/variances."""
import numpy as np
import pytest

from oracle import c_oracle, tgp_oracle as O
from tests.util import random_lgssm, relerr, sample_y, to_pkg_model

pytestmark = pytest.mark.gpu

LML_RTOL = 1e-6
MV_RTOL = 1e-5


def _check_filter(pkg, handle, m, y, atol=1e-9):
    pm = to_pkg_model(pkg, m)
    ms_o, Ps_o, lmls_o = O.filter_(m, y)
    lml, steps = pkg.lgssm.logpdf(pm, y, handle, per_step=True)
    assert abs(lml - lmls_o.sum()) <= LML_RTOL * abs(lmls_o.sum()) + 1e-9
    np.testing.assert_allclose(steps, lmls_o, rtol=1e-6, atol=1e-8)
    lml2 = pkg.lgssm.logpdf(pm, y, handle)
    assert abs(lml2 - lmls_o.sum()) <= LML_RTOL * abs(lmls_o.sum()) + 1e-9
    ms, Ps = pkg.lgssm._filter(pm, y, handle)
    np.testing.assert_allclose(ms, ms_o, rtol=MV_RTOL, atol=atol)
    np.testing.assert_allclose(Ps, Ps_o, rtol=MV_RTOL, atol=atol)


@pytest.mark.parametrize("algo", ["auto", "scan"])
@pytest.mark.parametrize("ordering", ["forward", "reverse"])
@pytest.mark.parametrize("tv", [True, False])
@pytest.mark.parametrize("D", [1, 2, 3, 4])
@pytest.mark.parametrize("T", [1, 2, 13, 49, 1000])
def test_filter_random_models(pkg, handle, T, D, tv, ordering, algo):
    handle.set_algo(pkg.TGP_ALGO_AUTO if algo == "auto" else pkg.TGP_ALGO_SCAN)
    rng = np.random.default_rng(1000 * T + 10 * D + tv)
    m = random_lgssm(rng, T, D, ordering, tv)
    y = sample_y(rng, m)
    try:
        _check_filter(pkg, handle, m, y)
    finally:
        handle.set_algo(pkg.TGP_ALGO_AUTO)


@pytest.mark.parametrize("chunk", [1, 3, 16, 64])
def test_chunk_sizes(pkg, handle, chunk):
    rng = np.random.default_rng(7)
    m = random_lgssm(rng, 5000, 3, "forward", True)
    y = sample_y(rng, m)
    handle.set_chunk(chunk)
    handle.set_algo(pkg.TGP_ALGO_SCAN)
    try:
        _check_filter(pkg, handle, m, y)
    finally:
        handle.set_chunk(0)
        handle.set_algo(pkg.TGP_ALGO_AUTO)


KERNELS = {
    "m12": (lambda p: p.Matern12Kernel(), lambda: O.Matern12()),
    "m32": (lambda p: p.Matern32Kernel(), lambda: O.Matern32()),
    "m52": (lambda p: p.Matern52Kernel(), lambda: O.Matern52()),
    "m52_scaled_stretched": (lambda p: 1.5 * p.with_lengthscale(p.Matern52Kernel(), 2.3),
                             lambda: 1.5 * O.Matern52().stretch(1 / 2.3)),
    "sum_m12_m32": (lambda p: p.Matern12Kernel() + 0.7 * p.Matern32Kernel(), lambda: O.Matern12() + 0.7 * O.Matern32()),
}


@pytest.mark.parametrize("regular", [True, False])
@pytest.mark.parametrize("kname", list(KERNELS))
def test_gp_logpdf_and_posterior_small(pkg, kname, regular):
    """The reference's integration checks (test/gp/lti_sde.jl:192-201, posterior_lti_sde.jl:82-89)
    at its own size N=13, against the oracle AND the dense GP."""
    kp, ko = KERNELS[kname]
    N = 13
    rng = np.random.default_rng(123456)
    tp = pkg.RegularSpacing(0.0, 0.3, N) if regular else pkg.RegularSpacing(0.0, 0.3, N).collect()
    to = O.RegularSpacing(0.0, 0.3, N) if regular else O.RegularSpacing(0.0, 0.3, N).collect()
    y = O.sample_prior(O.build_lgssm(ko(), to, 0.1), rng)
    fx = pkg.to_sde(pkg.GP(kp(pkg)))(tp, 0.1)
    lml = pkg.gp.logpdf(fx, y)
    assert abs(lml - O.gp_logpdf(ko(), to, 0.1, y)) <= LML_RTOL * abs(lml)
    assert abs(lml - O.dense_logpdf(ko(), to, 0.1, y)) <= 1e-6 * abs(lml)
    mu, var = pkg.gp.marginals(fx)
    mu_d, var_d = O.dense_prior_marginals(ko(), to, 0.1)
    np.testing.assert_allclose(mu, mu_d, atol=1e-9)
    np.testing.assert_allclose(var, var_d, rtol=1e-9)
    # posterior at the training inputs and at new inputs
    post = pkg.gp.posterior(fx, y)
    mu, var = pkg.gp.marginals(post(tp, 0.3))
    mu_o, var_o = O.gp_posterior_marginals(ko(), to, 0.1, y, None, 0.3)
    np.testing.assert_allclose(mu, mu_o, rtol=MV_RTOL, atol=1e-8)
    np.testing.assert_allclose(var, var_o, rtol=MV_RTOL)
    t_pr = np.sort(rng.uniform(-0.5, 4.5, 5))
    mu, var = pkg.gp.marginals(post(t_pr, 0.3))
    mu_d, cov_d = O.dense_posterior(ko(), to, 0.1, y, t_pr, 0.3)
    np.testing.assert_allclose(mu, mu_d, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(var, np.diag(cov_d), rtol=1e-5)
    y_pr = rng.standard_normal(5)
    lp = pkg.gp.logpdf(post(t_pr, 0.3), y_pr)
    lp_d = O.dense_posterior_logpdf(ko(), to, 0.1, y, t_pr, 0.3, y_pr)
    assert abs(lp - lp_d) <= 1e-5 * abs(lp_d)


@pytest.mark.parametrize("T", [1, 5, 257, 4000])
@pytest.mark.parametrize("D", [1, 2, 3])
def test_posterior_marginals_and_dynamics(pkg, handle, T, D):
    rng = np.random.default_rng(50 + T + D)
    m = random_lgssm(rng, T, D, "forward", True)
    y = sample_y(rng, m)
    pm = to_pkg_model(pkg, m)
    Rn = rng.uniform(0.01, 0.5, T)
    post_o = O.posterior(m, y)
    mu_o, var_o = O.marginals(O.replace_observation_noise_cov(post_o, Rn))
    mu, var, lml = pkg.lgssm.posterior_marginals(pm, y, Rn, handle, return_lml=True)
    np.testing.assert_allclose(mu, mu_o, rtol=MV_RTOL, atol=1e-8)
    np.testing.assert_allclose(var, var_o, rtol=MV_RTOL)
    assert abs(lml - O.logpdf(m, y)) <= LML_RTOL * abs(lml) + 1e-9
    post = pkg.lgssm.posterior(pm, y, handle)
    np.testing.assert_allclose(post.transitions.As, post_o.As, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(post.transitions.as_, post_o.as_, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(post.transitions.Qs, post_o.Qs, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(post.transitions.x0.m, post_o.m0, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(post.transitions.x0.P, post_o.P0, rtol=1e-6, atol=1e-9)
    # marginals of the materialised Reverse model run through tgp_marginals
    mu2, var2 = pkg.lgssm.marginals(pkg.lgssm.replace_observation_noise_cov(post, Rn), handle)
    np.testing.assert_allclose(mu2, mu_o, rtol=MV_RTOL, atol=1e-8)
    np.testing.assert_allclose(var2, var_o, rtol=MV_RTOL)


def test_missing_observations(pkg, handle):
    """test/models/missings.jl:94-115 — missing == NaN; compared with the oracle's transform."""
    rng = np.random.default_rng(9)
    m = random_lgssm(rng, 300, 3, "forward", True)
    y = sample_y(rng, m)
    y[rng.uniform(size=300) < 0.3] = np.nan
    pm = to_pkg_model(pkg, m)
    lml = pkg.lgssm.logpdf(pm, np.ma.masked_invalid(y), handle)
    ref = O.logpdf_missing(m, y)
    assert abs(lml - ref) <= LML_RTOL * abs(ref)


def test_cfg1_matern32_readme(pkg, handle):
    """BASELINE config 1: Matern32, RegularSpacing(0, 0.1, 10_000), sigma^2 = 0.1 (README.md:28-42)."""
    T = 10_000
    to = O.RegularSpacing(0.0, 0.1, T)
    mo = O.build_lgssm(O.Matern32(), to, 0.1)
    y = O.sample_prior(mo, np.random.default_rng(20261017 + 1))
    cm = c_oracle.Model.from_lgssm(mo)
    ref = c_oracle.filter(cm, y)
    fx = pkg.to_sde(pkg.GP(pkg.Matern32Kernel()), pkg.SArrayStorage(np.float64))(pkg.RegularSpacing(0.0, 0.1, T), 0.1)
    for algo in (pkg.TGP_ALGO_AUTO, pkg.TGP_ALGO_SCAN):
        handle.set_algo(algo)
        try:
            lml = pkg.gp.logpdf(fx, y)
            assert abs(lml - ref["lml"]) <= LML_RTOL * abs(ref["lml"])
            ms, Ps = pkg.lgssm._filter(fx.build_lgssm(), y, handle)
            np.testing.assert_allclose(ms, ref["m"], rtol=MV_RTOL, atol=1e-9)
            np.testing.assert_allclose(Ps, ref["P"], rtol=MV_RTOL, atol=1e-12)
        finally:
            handle.set_algo(pkg.TGP_ALGO_AUTO)
    # first 2 000 points against the dense GP (SURVEY.md §8d cfg 1)
    n = 2000
    fx2 = pkg.to_sde(pkg.GP(pkg.Matern32Kernel()))(pkg.RegularSpacing(0.0, 0.1, n), 0.1)
    d = O.dense_logpdf(O.Matern32(), O.RegularSpacing(0.0, 0.1, n), 0.1, y[:n])
    assert abs(pkg.gp.logpdf(fx2, y[:n]) - d) <= 1e-6 * abs(d)
    mu, var = pkg.gp.marginals(pkg.gp.posterior(fx, y)(pkg.RegularSpacing(0.0, 0.1, T), 1e-2))
    mu_o, var_o, _ = c_oracle.posterior_marginals(cm, y, 1e-2)
    np.testing.assert_allclose(mu, mu_o, rtol=MV_RTOL, atol=1e-8)
    np.testing.assert_allclose(var, var_o, rtol=MV_RTOL)


@pytest.mark.parametrize("T", [10**6])
def test_cfg2_matern52_large(pkg, handle, T):
    """BASELINE config 2 shape (Matern52, dt = 0.01, sigma^2 = 0.1) at a size the C oracle runs in
    well under a second."""
    to = O.RegularSpacing(0.0, 0.01, T)
    mo = O.build_lgssm(O.Matern52(), to, 0.1)
    rng = np.random.default_rng(20261017 + 2)
    y = np.cumsum(rng.standard_normal(T)) * 0.01 + rng.standard_normal(T) * 0.3
    cm = c_oracle.Model.from_lgssm(mo)
    ref = c_oracle.filter(cm, y)
    fx = pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 0.01, T), 0.1)
    for algo in (pkg.TGP_ALGO_AUTO, pkg.TGP_ALGO_SCAN):
        handle.set_algo(algo)
        try:
            lml, steps = pkg.lgssm.logpdf(fx.build_lgssm(), y, handle, per_step=True)
            assert abs(lml - ref["lml"]) <= LML_RTOL * abs(ref["lml"])
            np.testing.assert_allclose(steps, ref["lml_steps"], rtol=1e-6, atol=1e-8)
            ms, Ps = pkg.lgssm._filter(fx.build_lgssm(), y, handle)
            np.testing.assert_allclose(ms, ref["m"], rtol=MV_RTOL, atol=1e-8)
            np.testing.assert_allclose(Ps, ref["P"], rtol=MV_RTOL, atol=1e-12)
        finally:
            handle.set_algo(pkg.TGP_ALGO_AUTO)


def test_errors(pkg, handle):
    rng = np.random.default_rng(3)
    m = random_lgssm(rng, 20, 2, "forward", True)
    pm = to_pkg_model(pkg, m)
    with pytest.raises(pkg.DimensionMismatch):
        pkg.lgssm.logpdf(pm, np.zeros(19), handle)
    # negative innovation variance -> PosDefException-like status with the failing index
    bad = O.LGSSM("forward", m.As, m.as_, m.Qs, m.m0, m.P0, m.Hs, m.hs, np.full(20, -1e6))
    with pytest.raises(pkg.PosDefException):
        pkg.lgssm.logpdf(to_pkg_model(pkg, bad), np.zeros(20), handle)
    big = random_lgssm(rng, 5, 7, "forward", True)     # D = 7 has no scan instantiation: runs the dense path, still correct
    yb = sample_y(rng, big)
    assert abs(pkg.lgssm.logpdf(to_pkg_model(pkg, big), yb, handle) - O.logpdf(big, yb)) <= 1e-9 * abs(O.logpdf(big, yb))
    # D = 7 through the smoother entry points: the step-by-step path (tgp_seq.cu)
    mu, var = pkg.lgssm.posterior_marginals(to_pkg_model(pkg, big), yb, np.full(5, 0.1), handle)
    mu_o, var_o = O.marginals(O.replace_observation_noise_cov(O.posterior(big, yb), np.full(5, 0.1)))
    np.testing.assert_allclose(mu, mu_o, rtol=MV_RTOL, atol=1e-8)
    np.testing.assert_allclose(var, var_o, rtol=MV_RTOL)
    huge = random_lgssm(rng, 3, 2, "forward", True)
    huge_pm = to_pkg_model(pkg, huge)
    huge_pm.transitions.x0 = pkg.lgssm.Gaussian(np.zeros(70), np.eye(70))      # D = 70 > 64: refused, not mis-computed
    with pytest.raises(pkg.TGPError):
        pkg.lgssm.marginals(pkg.lgssm.LGSSM(pkg.lgssm.GaussMarkovModel("forward", np.stack([np.eye(70)] * 3), np.zeros((3, 70)),
                                                                       np.stack([np.eye(70)] * 3), huge_pm.transitions.x0),
                                            pkg.lgssm.ScalarEmissions(np.ones((3, 70)), np.zeros(3), np.ones(3))), handle)


def test_shard_reduce_prefix(pkg, handle):
    """Time-sharded logpdf on one GPU: 4 shards, elements folded on the host."""
    rng = np.random.default_rng(11)
    T, D, G = 4000, 3, 4
    m = random_lgssm(rng, T, D, "forward", True)
    y = sample_y(rng, m)
    ref = O.logpdf(m, y)
    ES = 3 * D * D + 2 * D
    elems = np.zeros((G, ES))
    bounds = np.linspace(0, T, G + 1).astype(int)
    shards = []
    for r in range(G):
        s, e = bounds[r], bounds[r + 1]
        sm = O.LGSSM("forward", m.As[s:e], m.as_[s:e], m.Qs[s:e], m.m0, m.P0, m.Hs[s:e], m.hs[s:e], m.Rs[s:e])
        shards.append(sm)
        mm = pkg.lgssm._Marshalled(to_pkg_model(pkg, sm))
        handle.shard_reduce(mm.desc, np.ascontiguousarray(y[s:e]), elems[r])
    total = 0.0
    for r in range(G):
        s, e = bounds[r], bounds[r + 1]
        m_in, P_in = handle.shard_prefix(D, elems[:r] if r else None, m.m0, m.P0)
        sm = shards[r]
        sm2 = O.LGSSM("forward", sm.As, sm.as_, sm.Qs, m_in, P_in, sm.Hs, sm.hs, sm.Rs)
        total += pkg.lgssm.logpdf(to_pkg_model(pkg, sm2), y[s:e], handle)
    assert abs(total - ref) <= LML_RTOL * abs(ref)


@pytest.mark.parametrize("T", [65536, 65537, 70001, 300_000])
@pytest.mark.parametrize("kname", ["m12", "m32", "m52", "sum_m12_m32"])
def test_steady_state_path_edges(pkg, handle, kname, T):
    """Time-invariant models long enough for the steady-state kernels: ragged tails, every output kind,
    strided (Julia Vector{Gaussian}-style) filter records; checked against the C oracle."""
    kp, ko = KERNELS[kname]
    to = O.RegularSpacing(0.0, 0.05, T)
    mo = O.build_lgssm(ko(), to, 0.2, 0.7)
    rng = np.random.default_rng(T)
    y = np.sin(np.arange(T) * 0.003) + 0.7 + 0.4 * rng.standard_normal(T)
    cm = c_oracle.Model.from_lgssm(mo)
    ref = c_oracle.filter(cm, y)
    fx = pkg.to_sde(pkg.GP(kp(pkg), 0.7))(pkg.RegularSpacing(0.0, 0.05, T), 0.2)
    model = fx.build_lgssm()
    c0 = handle.counters()["launches"]
    lml, steps = pkg.lgssm.logpdf(model, y, handle, per_step=True)
    assert handle.counters()["launches"] - c0 == 2, "expected the two-launch steady-state path"
    assert abs(lml - ref["lml"]) <= LML_RTOL * abs(ref["lml"])
    np.testing.assert_allclose(steps, ref["lml_steps"], rtol=1e-6, atol=1e-8)
    ms, Ps = pkg.lgssm._filter(model, y, handle)
    np.testing.assert_allclose(ms, ref["m"], rtol=MV_RTOL, atol=1e-8)
    np.testing.assert_allclose(Ps, ref["P"], rtol=MV_RTOL, atol=1e-12)
    # strided records: m at +0, P at +D of each (D + D*D)-double record
    D = mo.D
    rec = np.zeros((T, D + D * D))
    mm = pkg.lgssm._Marshalled(model)
    out = np.zeros(1)
    base = rec.ctypes.data
    handle.filter(mm.desc, np.ascontiguousarray(y), base, D + D * D, base + 8 * D, D + D * D, out)
    np.testing.assert_allclose(rec[:, :D], ref["m"], rtol=MV_RTOL, atol=1e-8)
    np.testing.assert_allclose(rec[:, D:].reshape(T, D, D), np.swapaxes(ref["P"], 1, 2), rtol=MV_RTOL, atol=1e-12)
    assert abs(out[0] - ref["lml"]) <= LML_RTOL * abs(ref["lml"])


def test_steady_state_fallback_when_not_converged(pkg, handle):
    """A model whose covariance has not converged within the transient budget (dt tiny against the length
    scale) must silently rerun through the general scan and still match."""
    T = 100_000
    mo = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 1e-6, T), 0.5)
    rng = np.random.default_rng(5)
    y = 0.3 * rng.standard_normal(T)
    ref = c_oracle.filter(c_oracle.Model.from_lgssm(mo), y, want_mp=False)
    fx = pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 1e-6, T), 0.5)
    c0 = handle.counters()["launches"]
    lml = pkg.lgssm.logpdf(fx.build_lgssm(), y, handle)
    assert handle.counters()["launches"] - c0 > 2   # steady attempt + general rerun
    assert abs(lml - ref["lml"]) <= LML_RTOL * abs(ref["lml"])


@pytest.mark.parametrize("world,T", [(2, 70_000), (3, 65_536 + 777), (4, 131_072), (8, 100_001)])
def test_steady_sharded_phases_single_device(pkg, world, T):
    """The two-phase steady-state sharded logpdf (tgp_shard_phase1 / tgp_shard_phase2) with every 'rank' played by its
    own handle on cuda:0 and the all-gather done by hand: exercises the exact end-of-shard alignment (ragged last
    tile) and the record fold; compared with the sequential C oracle on the whole series."""
    import torch
    dev = torch.device("cuda:0")
    Ttot = T * world
    mo = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.02, Ttot), 0.3, 0.4)
    rng = np.random.default_rng(world * 1000 + T)
    y = np.cos(np.arange(Ttot) * 0.001) + 0.4 + 0.5 * rng.standard_normal(Ttot)
    ref = c_oracle.logpdf(c_oracle.Model.from_lgssm(mo), y)
    fx = pkg.to_sde(pkg.GP(pkg.Matern52Kernel(), 0.4))(pkg.RegularSpacing(0.0, 0.02, T), 0.3)
    mm = pkg.lgssm._Marshalled(fx.build_lgssm())
    D = mm.D
    XS = D * D + D
    handles = [pkg.Handle(0) for _ in range(world)]
    ys = [torch.from_numpy(np.ascontiguousarray(y[r * T:(r + 1) * T])).to(dev) for r in range(world)]
    recs = torch.zeros(world * XS, dtype=torch.float64, device=dev)
    for r in range(world):
        handles[r].shard_phase1(mm.desc, ys[r], r, world, recs[r * XS:(r + 1) * XS])
    parts = torch.zeros(world, dtype=torch.float64, device=dev)
    for r in range(world):
        handles[r].shard_phase2(recs, parts[r:r + 1])
    for r in range(world):
        handles[r].synchronize()
    total = float(parts.sum().item())
    assert abs(total - ref) <= LML_RTOL * abs(ref), (total, ref)
    # the records chain to the sequential filter's mean at every shard boundary
    ms = c_oracle.filter(c_oracle.Model.from_lgssm(mo), y, want_steps=False)["m"]
    R = recs.cpu().numpy().reshape(world, XS)
    x = np.zeros(D)
    for r in range(world - 1):
        Phi = R[r, :D * D].reshape(D, D).T
        x = Phi @ x + R[r, D * D:]
        np.testing.assert_allclose(x, ms[(r + 1) * T - 1], rtol=1e-6, atol=1e-8)
    for h in handles:
        h.close()


@pytest.mark.parametrize("ordering", ["forward", "reverse"])
@pytest.mark.parametrize("D", [5, 6, 8, 10])
def test_filter_larger_state_dims(pkg, handle, D, ordering):
    """D = 8 and 10 run the rolled-loop instantiations (state in thread-local memory)."""
    rng = np.random.default_rng(D)
    m = random_lgssm(rng, 700, D, ordering, True)
    y = sample_y(rng, m)
    _check_filter(pkg, handle, m, y)


def cfg3_kernels(pkg):
    """BASELINE config 3: D = 10 sum  Matern32 + 0.7 Matern52 + 0.5 Matern52∘ST(0.5) + 0.3 Matern32∘ST(2.0)
    (the reference has no RQ kernel; this is a reference-expressible stand-in of the same state dimension)."""
    TK = pkg.gp.TransformedKernel
    kp = (1.0 * pkg.Matern32Kernel() + 0.7 * pkg.Matern52Kernel() + 0.5 * TK(pkg.Matern52Kernel(), 0.5)
          + 0.3 * TK(pkg.Matern32Kernel(), 2.0))
    ko = 1.0 * O.Matern32() + 0.7 * O.Matern52() + 0.5 * O.Matern52().stretch(0.5) + 0.3 * O.Matern32().stretch(2.0)
    return kp, ko


@pytest.mark.parametrize("T", [3000, 50_000])
def test_cfg3_sum_kernel_d10_posterior_marginals(pkg, handle, T):
    kp, ko = cfg3_kernels(pkg)
    to = O.RegularSpacing(0.0, 0.01, T)
    mo = O.build_lgssm(ko, to, 0.1)
    assert mo.D == 10
    rng = np.random.default_rng(20261017 + 3)
    y = np.sin(np.arange(T) * 0.004) + 0.3 * np.cos(np.arange(T) * 0.05) + 0.35 * rng.standard_normal(T)
    cm = c_oracle.Model.from_lgssm(mo)
    mu_o, var_o, lml_o = c_oracle.posterior_marginals(cm, y, 1e-2)
    fx = pkg.to_sde(pkg.GP(kp))(pkg.RegularSpacing(0.0, 0.01, T), 0.1)
    lml = pkg.gp.logpdf(fx, y)
    assert abs(lml - lml_o) <= LML_RTOL * abs(lml_o)
    mu, var = pkg.gp.marginals(pkg.gp.posterior(fx, y)(pkg.RegularSpacing(0.0, 0.01, T), 1e-2))
    np.testing.assert_allclose(mu, mu_o, rtol=MV_RTOL, atol=1e-7)
    np.testing.assert_allclose(var, var_o, rtol=MV_RTOL)


# ---- vector observations / large state: the dense step-by-step path (tgp_dense.cu) ----------------------------------
def _random_vector_lgssm(rng, T, D, M, ordering, r_dense):
    from tests.util import random_psd, _stable
    As = np.stack([_stable(0.9 * np.eye(D) + 0.1 * rng.standard_normal((D, D))) for _ in range(T)])
    as_ = rng.standard_normal((T, D)) * 0.3
    Qs = np.stack([random_psd(rng, D) for _ in range(T)])
    Hs = rng.standard_normal((T, M, D))
    hs = rng.standard_normal((T, M)) * 0.2
    Rs = np.stack([random_psd(rng, M, 0.1, 1.0) if r_dense else np.diag(rng.uniform(0.1, 1.0, M)) for _ in range(T)])
    return O.LGSSM(ordering, As, as_, Qs, rng.standard_normal(D), random_psd(rng, D, 0.5, 2.0), Hs, hs, Rs)


def _pkg_vector_model(pkg, m, r_dense):
    L = pkg.lgssm
    tr = L.GaussMarkovModel(m.ordering, np.array(m.As), np.array(m.as_), np.array(m.Qs), L.Gaussian(np.array(m.m0), np.array(m.P0)))
    Rs = np.array(m.Rs) if r_dense else np.array([np.diag(R) for R in m.Rs])
    return L.LGSSM(tr, L.SmallOutputEmissions(np.array(m.Hs), np.array(m.hs), Rs))


@pytest.mark.parametrize("r_dense", [False, True])
@pytest.mark.parametrize("ordering", ["forward", "reverse"])
@pytest.mark.parametrize("D,M", [(1, 1), (3, 2), (3, 1), (7, 3), (12, 5)])
def test_small_output_lgc_vector_observations(pkg, handle, D, M, ordering, r_dense):
    """posterior_and_lml(::SmallOutputLGC) (LGC:129-141) through tgp_logpdf / tgp_filter: Dlat/Dobs of the reference's
    unit tests (test/models/lgssm.jl: Dlat in {1,3}, Dobs in {1,2}) and larger; time-varying, both orderings."""
    rng = np.random.default_rng(100 * D + M)
    T = 49
    m = _random_vector_lgssm(rng, T, D, M, ordering, r_dense)
    y = O.sample_prior(O.LGSSM("forward", m.As, m.as_, m.Qs, m.m0, m.P0, m.Hs, m.hs, m.Rs), rng)
    ms_o, Ps_o, lmls_o = O.filter_(m, y)
    pm = _pkg_vector_model(pkg, m, r_dense)
    lml, steps = pkg.lgssm.logpdf(pm, y, handle, per_step=True)
    np.testing.assert_allclose(steps, lmls_o, rtol=1e-9, atol=1e-10)
    assert abs(lml - lmls_o.sum()) <= LML_RTOL * abs(lmls_o.sum())
    ms, Ps = pkg.lgssm._filter(pm, y, handle)
    np.testing.assert_allclose(ms, ms_o, rtol=MV_RTOL, atol=1e-9)
    np.testing.assert_allclose(Ps, Ps_o, rtol=MV_RTOL, atol=1e-10)


@pytest.mark.parametrize("regular", [True, False])
def test_space_time_separable_logpdf(pkg, regular):
    """test/space_time/to_gauss_markov.jl:36-66: Separable(SE, Matern32) on RectilinearGrid(Nr = 3, Nt = 5): SDE path ==
    dense GP; plus a larger grid against the oracle (time-invariant -> the CUDA-graph replay)."""
    rng = np.random.default_rng(123456)
    r = rng.standard_normal(3)
    tp = pkg.RegularSpacing(0.0, 0.3, 5) if regular else np.sort(rng.uniform(0, 2, 5))
    to = O.RegularSpacing(0.0, 0.3, 5) if regular else np.array(tp)
    mo = O.build_lgssm_separable(O.SqExp(), O.Matern32(), r, to, 0.1)
    y = O.sample_prior(mo, rng)
    fx = pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern32Kernel())))(pkg.RectilinearGrid(r, tp), 0.1)
    lml = pkg.gp.logpdf(fx, y.reshape(-1))
    lp_d = O.dense_separable_logpdf(O.SqExp(), O.Matern32(), r, to, 0.1, y.reshape(-1))
    assert abs(lml - lp_d) <= 1e-6 * abs(lp_d)
    assert abs(lml - O.logpdf(mo, y)) <= LML_RTOL * abs(lml)
    # config-5 shape, scaled down: Separable(SE, Matern52), 24 spatial points x 300 times (D = 72, M = 24)
    r2 = np.linspace(-3, 3, 24)
    T = 300
    mo2 = O.build_lgssm_separable(O.SqExp(), O.Matern52(), r2, O.RegularSpacing(0.0, 0.01, T), 0.1)
    y2 = O.sample_prior(mo2, rng)
    fx2 = pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern52Kernel())))(pkg.RectilinearGrid(r2, pkg.RegularSpacing(0.0, 0.01, T)), 0.1)
    lml2 = pkg.gp.logpdf(fx2, y2.reshape(-1))
    ref2 = O.logpdf(mo2, y2)
    assert abs(lml2 - ref2) <= LML_RTOL * abs(ref2)


# ---- small observation noise (the reference's DEFAULT is 1e-12, lti_sde.jl:27-29) and non-contractive dynamics -------------------
@pytest.mark.parametrize("regular", [True, False])
@pytest.mark.parametrize("s2", [1e-6, 1e-9, 1e-12])
@pytest.mark.parametrize("kname", ["m52", "m32", "sum_m12_m32"])
def test_small_noise_parity(pkg, handle, kname, s2, regular, monkeypatch):
    """logpdf (1e-6), filtering means / covariances and posterior marginals (1e-5) at observation noise down to the reference's
    default 1e-12, through the steady-state kernels (regular grid) and the general 5-tuple scan (irregular grid). Covariance entries
    are compared relative to the largest entry (a state observed with noise 1e-12 has variances spanning 12 decades)."""
    kp, ko = KERNELS[kname]
    # Same transitions, bit for bit, on both sides (host-built): at sigma^2 = 1e-12 the result is sensitive to the LAST bits of
    # Q_t = P - A_t P A_t' (catastrophic cancellation for small gaps), so an independently rounded exp(F dt) — the device builder's,
    # or Julia's — moves the answer by ~1e-5 relative. tests/test_gpu_lti.py measures exactly that for the device builder.
    monkeypatch.setattr(pkg.gp, "DEVICE_COMPONENTS_MIN_T", 10 ** 9)
    rng = np.random.default_rng(int(-np.log10(s2)) * 10 + regular)
    T, dt = 6000, 0.01
    tp = pkg.RegularSpacing(0.0, dt, T) if regular else np.sort(rng.uniform(0, dt * T, T))
    to = O.RegularSpacing(0.0, dt, T) if regular else np.array(tp)
    y = O.sample_prior(O.build_lgssm(ko(), to, 0.1), rng)
    mo = O.build_lgssm(ko(), to, s2)
    fx = pkg.to_sde(pkg.GP(kp(pkg)))(tp, s2)
    ref = O.logpdf(mo, y)
    lml = pkg.gp.logpdf(fx, y)
    assert abs(lml - ref) <= LML_RTOL * abs(ref), (lml, ref)
    ms_o, Ps_o, _ = O.filter_(mo, y)
    ms, Ps = pkg.lgssm._filter(fx.build_lgssm(), y, handle)
    np.testing.assert_allclose(ms, ms_o, rtol=MV_RTOL, atol=1e-7)
    np.testing.assert_allclose(Ps, Ps_o, rtol=MV_RTOL, atol=1e-9 * np.abs(Ps_o).max())
    mu, var = pkg.gp.marginals(pkg.gp.posterior(fx, y)(tp, 1e-2))
    mu_o, var_o = O.gp_posterior_marginals(ko(), to, s2, y, None, 1e-2)
    np.testing.assert_allclose(mu, mu_o, rtol=MV_RTOL, atol=1e-7)
    np.testing.assert_allclose(var, var_o, rtol=MV_RTOL)


def test_non_contractive_dynamics_are_right_or_loud(pkg, handle):
    """The reference's fixtures use A = I + 0.1 randn (test/models/model_test_utils.jl:29-31), which is not contractive. At its own
    lengths (N <= 49) the scan must simply agree; on a long horizon the intermediate scan elements of an UNSTABLE model overflow —
    then the call has to fail loudly (TGP_ENOTPD from the non-finite innovation variance), never return a wrong number."""
    rng = np.random.default_rng(99)
    D = 3
    for T in (49, 400, 4000):
        n = T
        As = np.stack([np.eye(D) + 0.1 * rng.standard_normal((D, D)) for _ in range(n)])
        from tests.util import random_psd
        m = O.LGSSM("forward", As, rng.standard_normal((n, D)) * 0.3, np.stack([random_psd(rng, D) for _ in range(n)]),
                    rng.standard_normal(D), random_psd(rng, D, 0.5, 2.0), rng.standard_normal((n, D)), rng.standard_normal(n) * 0.2,
                    rng.uniform(0.05, 1.0, n))
        y = rng.standard_normal(T)
        ref = O.logpdf(m, y)
        try:
            lml = pkg.lgssm.logpdf(to_pkg_model(pkg, m), y, handle)
        except pkg.TGPError:
            assert T > 49, "the reference's own fixture sizes must work"
            continue
        assert np.isfinite(ref) and abs(lml - ref) <= LML_RTOL * abs(ref), (T, lml, ref)
