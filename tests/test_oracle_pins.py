"""Pins the oracle (oracle/tgp_oracle.py and the C restatement) the way the reference pins itself:
every equivalence its own test-suite asserts for this path is re-executed against the oracle.
CPU only. The reference ships no golden vectors and cannot be run here (no Julia) — SURVEY.md §8c."""
import numpy as np
import pytest

from oracle import c_oracle, tgp_oracle as O
from tests.kernels import KERNEL_IDS, KERNELS, MEANS, N, inputs, noises
from tests.util import random_lgssm, sample_y


@pytest.mark.parametrize("regular", [True, False], ids=["regular", "irregular"])
@pytest.mark.parametrize("mean", MEANS, ids=[m[0] for m in MEANS])
@pytest.mark.parametrize("kernel", KERNELS, ids=KERNEL_IDS)
def test_sde_path_equals_dense_gp_prior(kernel, mean, regular):
    """test/gp/lti_sde.jl:192-201 — marginal means / variances and logpdf of the SDE path == naive GP."""
    _, ko, _ = kernel
    rng = np.random.default_rng(123456)
    t = inputs(regular, O)
    for _, s2 in noises(rng):
        model = O.build_lgssm(ko(), t, s2, mean[1])
        y = O.sample_prior(model, rng)
        mu, var = O.gp_prior_marginals(ko(), t, s2, mean[1])
        mu_d, var_d = O.dense_prior_marginals(ko(), t, s2, mean[1])
        np.testing.assert_allclose(mu, mu_d, rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(var, var_d, rtol=1e-7)
        lp, lp_d = O.gp_logpdf(ko(), t, s2, y, mean[1]), O.dense_logpdf(ko(), t, s2, y, mean[1])
        assert abs(lp - lp_d) <= 1.5e-8 * abs(lp_d) + 1e-8   # default `≈`: rtol sqrt(eps)


@pytest.mark.parametrize("kernel", [KERNELS[i] for i in (0, 1, 2, 5, 9, 16, 17)], ids=[KERNEL_IDS[i] for i in (0, 1, 2, 5, 9, 16, 17)])
def test_sde_posterior_equals_dense_gp_posterior(kernel):
    """test/gp/posterior_lti_sde.jl:52-90 — N = 3 training / 2 prediction points, predict noise 0.3, rtol 1e-5;
    and at larger sizes."""
    _, ko, _ = kernel
    for n, n_pr in ((3, 2), (13, 7)):
        rng = np.random.default_rng(123456 + n)
        t = np.sort(rng.uniform(0.0, 3.0, n))
        t_pr = np.sort(rng.uniform(-0.5, 3.5, n_pr))
        y = rng.standard_normal(n)
        mu, var = O.gp_posterior_marginals(ko(), t, 0.1, y, t_pr, 0.3)
        mu_d, cov_d = O.dense_posterior(ko(), t, 0.1, y, t_pr, 0.3)
        np.testing.assert_allclose(mu, mu_d, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(var, np.diag(cov_d), rtol=1e-5)
        y_pr = rng.standard_normal(n_pr)
        lp = O.gp_posterior_logpdf(ko(), t, 0.1, y, t_pr, 0.3, y_pr)
        lp_d = O.dense_posterior_logpdf(ko(), t, 0.1, y, t_pr, 0.3, y_pr)
        assert abs(lp - lp_d) <= 1e-5 * abs(lp_d)
        # same inputs branch (posterior_lti_sde.jl:27-36)
        mu, var = O.gp_posterior_marginals(ko(), t, 0.1, y, None, 0.3)
        mu_d, cov_d = O.dense_posterior(ko(), t, 0.1, y, t, 0.3)
        np.testing.assert_allclose(mu, mu_d, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(var, np.diag(cov_d), rtol=1e-5)


def test_space_time_separable_equals_dense():
    """test/space_time/to_gauss_markov.jl:36-66 — Separable(SE, Matern32) on a RectilinearGrid(Nr = 3, Nt = 5)."""
    rng = np.random.default_rng(123456)
    r = rng.standard_normal(3)
    for t in (O.RegularSpacing(0.0, 0.3, 5), np.sort(rng.uniform(0, 2, 5))):
        model = O.build_lgssm_separable(O.SqExp(), O.Matern32(), r, t, 0.1)
        y = O.sample_prior(model, rng)
        lp = O.logpdf(model, y)
        lp_d = O.dense_separable_logpdf(O.SqExp(), O.Matern32(), r, t, 0.1, y.reshape(-1))
        assert abs(lp - lp_d) <= 1e-6 * abs(lp_d)


def test_missing_equals_marginalised_model():
    """test/models/missings.jl:94-115 — a missing observation == the analytically marginalised model: here the
    dense GP with the missing points deleted."""
    rng = np.random.default_rng(3)
    k = O.Matern52()
    t = np.sort(rng.uniform(0, 4, 12))
    y = rng.standard_normal(12)
    miss = np.zeros(12, bool)
    miss[[2, 3, 7]] = True
    ym = y.copy()
    ym[miss] = np.nan
    lp = O.gp_logpdf(k, t, 0.2, ym)
    lp_d = O.dense_logpdf(k, t[~miss], 0.2, y[~miss])
    assert abs(lp - lp_d) <= 1e-8 * abs(lp_d) + 1e-8   # lml atol/rtol 1e-8 in the reference


@pytest.mark.parametrize("D", [1, 3])
def test_scalar_lgc_equals_1x1_small_lgc(D):
    """test/models/linear_gaussian_conditionals.jl:117-126; LargeOutputLGC == SmallOutputLGC :65-75."""
    rng = np.random.default_rng(D)
    m = rng.standard_normal(D)
    A = rng.standard_normal((D, D))
    P = A @ A.T + 0.1 * np.eye(D)
    H = rng.standard_normal(D)
    h, R, y = 0.3, 0.7, 1.1
    m1, P1, l1 = O.posterior_and_lml_scalar(m, P, H, h, R, y)
    m2, P2, l2 = O.posterior_and_lml_small(m, P, H[None, :], np.array([h]), np.array([[R]]), np.array([y]))
    m3, P3, l3 = O.posterior_and_lml_large(m, P, H[None, :], np.array([h]), np.array([[R]]), np.array([y]))
    for a, b in ((m1, m2), (P1, P2), (l1, l2)):
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)
    for a, b in ((m3, m2), (P3, P2), (l3, l2)):
        np.testing.assert_allclose(a, b, rtol=1.5e-8, atol=1e-8)


@pytest.mark.parametrize("ordering", ["forward", "reverse"])
@pytest.mark.parametrize("tv", [True, False])
@pytest.mark.parametrize("D", [1, 2, 3, 4, 6])
def test_c_restatement_equals_numpy_restatement(D, tv, ordering):
    """The two independent restatements (NumPy, C; static and run-time sized) agree to rounding on every output."""
    rng = np.random.default_rng(17 * D + tv)
    m = random_lgssm(rng, 49, D, ordering, tv)
    y = sample_y(rng, m)
    ms, Ps, lmls = O.filter_(m, y)
    for static in (True, False):
        c_oracle.set_static(static)
        r = c_oracle.filter(c_oracle.Model.from_lgssm(m), y)
        np.testing.assert_allclose(r["m"], ms, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(r["P"], Ps, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(r["lml_steps"], lmls, rtol=1e-11, atol=1e-12)
        assert abs(r["lml"] - O.logpdf(m, y)) <= 1e-12 * abs(r["lml"])
        mu, var = c_oracle.marginals(c_oracle.Model.from_lgssm(m))
        mu_o, var_o = O.marginals(m)
        np.testing.assert_allclose(mu, mu_o, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(var, var_o, rtol=1e-11)
        if ordering == "forward":
            po = O.posterior(m, y)
            pc = c_oracle.posterior(c_oracle.Model.from_lgssm(m), y)
            np.testing.assert_allclose(pc["G"], po.As, rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(pc["Sig"], po.Qs, rtol=1e-7, atol=1e-10)
            Rn = rng.uniform(0.01, 0.3, 49)
            mu, var, _ = c_oracle.posterior_marginals(c_oracle.Model.from_lgssm(m), y, Rn)
            mu_o, var_o = O.marginals(O.replace_observation_noise_cov(po, Rn))
            np.testing.assert_allclose(mu, mu_o, rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(var, var_o, rtol=1e-8)
    c_oracle.set_static(True)
