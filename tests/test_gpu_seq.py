"""GPU parity tests of the smoother-side entry points outside the scan shapes (tgp_seq.cu): vector observations (SmallOutputLGC),
Reverse-ordered models, marginals_diag, missing observations (whole rows and single entries), and the space-time posterior of the
reference's own test (test/space_time/to_gauss_markov.jl:68-87) — against the sequential oracle and the dense GP.
Tolerances: north_star's 1e-6 (logpdf) / 1e-5 (means, variances)."""
import numpy as np
import pytest

from oracle import tgp_oracle as O
from tests.test_gpu_parity import _pkg_vector_model, _random_vector_lgssm
from tests.util import random_lgssm, sample_y, to_pkg_model

pytestmark = pytest.mark.gpu
MV = dict(rtol=1e-5, atol=1e-8)


def _sample(rng, m):
    return O.sample_prior(O.LGSSM("forward", m.As, m.as_, m.Qs, m.m0, m.P0, m.Hs, m.hs, m.Rs), rng)


@pytest.mark.parametrize("ordering", ["forward", "reverse"])
@pytest.mark.parametrize("D", [1, 2, 3])
def test_scalar_posterior_both_orderings(pkg, handle, D, ordering):
    """step_posterior(::Forward) and step_posterior(::Reverse) (lgssm.jl:215-228, the reference's unit grid runs both)."""
    rng = np.random.default_rng(10 * D + (ordering == "reverse"))
    T = 49
    m = random_lgssm(rng, T, D, ordering, True)
    y = sample_y(rng, m)
    post_o = O.posterior(m, y)
    post = pkg.lgssm.posterior(to_pkg_model(pkg, m), y, handle)
    assert post.ordering == post_o.ordering
    np.testing.assert_allclose(post.transitions.As, post_o.As, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(post.transitions.as_, post_o.as_, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(post.transitions.Qs, post_o.Qs, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(post.transitions.x0.m, post_o.m0, **MV)
    np.testing.assert_allclose(post.transitions.x0.P, post_o.P0, **MV)
    # marginals of the posterior (the reversed-ordering model) and its marginals_diag
    mu_o, var_o = O.marginals(post_o)
    mu, var = pkg.lgssm.marginals(post, handle)
    np.testing.assert_allclose(mu, mu_o, **MV)
    np.testing.assert_allclose(var, var_o, **MV)
    mu_d, var_d = pkg.lgssm.marginals_diag(post, handle)
    np.testing.assert_allclose(mu_d, mu_o, **MV)
    np.testing.assert_allclose(var_d, var_o, **MV)
    # the fused chain (posterior -> replace noise -> marginals) for a Reverse prior
    Rn = rng.uniform(0.01, 0.5, T)
    mu_o2, var_o2 = O.marginals(O.replace_observation_noise_cov(post_o, Rn))
    mu2, var2 = pkg.lgssm.posterior_marginals(to_pkg_model(pkg, m), y, Rn, handle)
    np.testing.assert_allclose(mu2, mu_o2, **MV)
    np.testing.assert_allclose(var2, var_o2, **MV)


@pytest.mark.parametrize("r_dense", [False, True])
@pytest.mark.parametrize("ordering", ["forward", "reverse"])
@pytest.mark.parametrize("D,M", [(1, 1), (3, 2), (3, 1), (7, 3), (12, 5)])
def test_vector_observation_posterior_and_marginals(pkg, handle, D, M, ordering, r_dense):
    """posterior / marginals / marginals_diag with SmallOutputLGC emissions (lgssm.jl:99-141, 193-238; Dlat / Dobs of the reference's
    unit tests and larger), time-varying, both orderings."""
    rng = np.random.default_rng(1000 + 100 * D + M)
    T = 29
    m = _random_vector_lgssm(rng, T, D, M, ordering, r_dense)
    y = _sample(rng, m)
    pm = _pkg_vector_model(pkg, m, r_dense)
    mu_o, cov_o = O.marginals(m)
    mu, cov = pkg.lgssm.marginals(pm, handle)
    np.testing.assert_allclose(mu, mu_o, **MV)
    np.testing.assert_allclose(cov, cov_o, **MV)
    mu_d, var_d = pkg.lgssm.marginals_diag(pm, handle)
    mu_do, var_do = O.marginals_diag(m)
    np.testing.assert_allclose(mu_d, mu_do, **MV)
    np.testing.assert_allclose(var_d, var_do, **MV)
    post_o = O.posterior(m, y)
    post = pkg.lgssm.posterior(pm, y, handle)
    np.testing.assert_allclose(post.transitions.As, post_o.As, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(post.transitions.as_, post_o.as_, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(post.transitions.Qs, post_o.Qs, rtol=1e-5, atol=1e-7)
    mu_po, cov_po = O.marginals(post_o)
    mu_p, cov_p = pkg.lgssm.marginals(post, handle)
    np.testing.assert_allclose(mu_p, mu_po, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(cov_p, cov_po, rtol=1e-5, atol=1e-7)
    # fused chain with a new (diagonal / dense) observation noise
    Rn = np.stack([np.diag(rng.uniform(0.05, 0.4, M)) for _ in range(T)])
    mu_fo, cov_fo = O.marginals(O.replace_observation_noise_cov(post_o, Rn))
    mu_f, var_f = pkg.lgssm.posterior_marginals(pm, y, Rn if r_dense else np.stack([np.diag(R) for R in Rn]), handle)
    np.testing.assert_allclose(mu_f, mu_fo, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(var_f, np.einsum("tii->ti", cov_fo), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("r_dense", [False, True])
@pytest.mark.parametrize("ordering", ["forward", "reverse"])
def test_vector_observations_missing_rows(pkg, handle, ordering, r_dense):
    """missings.jl:8-33, 59-74 with Dobs = 2 (test/models/missings.jl:94-115): whole observations missing. logpdf, filter and the
    posterior marginals equal the oracle's, and the model with the missing steps' data ignored."""
    rng = np.random.default_rng(77 + r_dense)
    T, D, M = 31, 3, 2
    m = _random_vector_lgssm(rng, T, D, M, ordering, r_dense)
    y = _sample(rng, m)
    miss = rng.random(T) < 0.3
    y_nan = y.copy()
    y_nan[miss] = np.nan
    pm = _pkg_vector_model(pkg, m, r_dense)
    ym = np.ma.masked_invalid(y_nan)
    lml = pkg.lgssm.logpdf(pm, ym, handle)
    ref = O.logpdf_missing(m, y_nan)
    assert abs(lml - ref) <= 1e-6 * abs(ref)
    m2, y2, _ = O.transform_model_and_obs(m, y_nan)
    ms_o, Ps_o, _ = O.filter_(m2, y2)
    ms, Ps = pkg.lgssm._filter(pm, ym, handle)
    np.testing.assert_allclose(ms, ms_o, **MV)
    np.testing.assert_allclose(Ps, Ps_o, **MV)
    post_o = O.posterior_missing(m, y_nan)
    mu_o, cov_o = O.marginals(post_o)
    mu, cov = pkg.lgssm.marginals(pkg.lgssm.posterior(pm, ym, handle), handle)
    np.testing.assert_allclose(mu, mu_o, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(cov, cov_o, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("ordering", ["forward", "reverse"])
def test_vector_observations_missing_entries(pkg, handle, ordering):
    """posterior_and_lml(x, ::SmallOutputLGC, y::AbstractVector{<:Union{Missing, <:Real}}) (LGC:143-151): single entries missing,
    Diagonal observation covariance. Dense covariance: refused, as the reference's MethodError."""
    rng = np.random.default_rng(5)
    T, D, M = 25, 4, 3
    m = _random_vector_lgssm(rng, T, D, M, ordering, False)
    y = _sample(rng, m)
    y_nan = y.copy()
    y_nan[rng.random((T, M)) < 0.25] = np.nan
    y_nan[3] = np.nan                      # and one whole row
    pm = _pkg_vector_model(pkg, m, False)
    ym = np.ma.masked_invalid(y_nan)
    lml = pkg.lgssm.logpdf(pm, ym, handle)
    ref = O.logpdf_missing(m, y_nan)
    assert abs(lml - ref) <= 1e-6 * abs(ref)
    # equivalent formulation: an entry that is missing carries no information — drop it from H, h, R (oracle, per step)
    m2, y2, _ = O.transform_model_and_obs(m, y_nan)
    ms_o, Ps_o, _ = O.filter_(m2, y2)
    ms, Ps = pkg.lgssm._filter(pm, ym, handle)
    np.testing.assert_allclose(ms, ms_o, **MV)
    np.testing.assert_allclose(Ps, Ps_o, **MV)
    md = _random_vector_lgssm(np.random.default_rng(6), T, D, M, ordering, True)
    with pytest.raises(pkg.TGPError):
        pkg.lgssm.logpdf(_pkg_vector_model(pkg, md, True), ym, handle)


@pytest.mark.parametrize("regular", [True, False])
def test_space_time_posterior_matches_dense_gp(pkg, regular):
    """test/space_time/to_gauss_markov.jl:36-87: Separable(SE, Matern32) on RectilinearGrid(Nr = 3, Nt = 5): prior marginals, and the
    posterior at the same locations and at different times (marginals and logpdf) against the dense GP."""
    rng = np.random.default_rng(123456)
    Nr, Nt, Nt_pr = 3, 5, 2
    r = rng.standard_normal(Nr)
    tp = pkg.RegularSpacing(0.0, 0.3, Nt) if regular else np.sort(rng.random(Nt))
    to = O.RegularSpacing(0.0, 0.3, Nt) if regular else np.array(tp)
    tt = to.collect() if regular else to
    ks, kt = O.SqExp(), O.Matern32()
    mo = O.build_lgssm_separable(ks, kt, r, to, 0.1)
    y = O.sample_prior(mo, rng).reshape(-1)
    f = pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern32Kernel())))
    x = pkg.RectilinearGrid(r, tp)
    fx = f(x, 0.1)
    # prior marginals == dense GP (k(x, x) + noise)
    mu, var = pkg.gp.marginals(fx)
    Kd = np.kron(O.kernelmatrix(kt, tt), O.kernelmatrix(ks, r))
    np.testing.assert_allclose(mu, 0.0, atol=1e-12)
    np.testing.assert_allclose(var, np.diag(Kd) + 0.1, rtol=1e-8)
    post = pkg.gp.posterior(fx, y)
    for t_pr in (tt, rng.standard_normal(Nt_pr)):
        x_pr = pkg.RectilinearGrid(r, tp if t_pr is tt else t_pr)
        mu_d, cov_d = O.dense_separable_posterior(ks, kt, r, tt, 0.1, y, t_pr, 0.1)
        order = np.argsort(t_pr, kind="stable")
        mu, var = pkg.gp.marginals(post(x_pr, 0.1))
        np.testing.assert_allclose(mu, mu_d, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(var, np.diag(cov_d), rtol=1e-5)
        y_post = mu_d + np.linalg.cholesky(cov_d) @ rng.standard_normal(len(mu_d))
        lp = pkg.gp.logpdf(post(x_pr, 0.1), y_post)
        Lc = np.linalg.cholesky(cov_d)
        z = np.linalg.solve(Lc, y_post - mu_d)
        lp_d = -0.5 * (len(mu_d) * np.log(2 * np.pi) + 2 * np.log(np.diag(Lc)).sum() + z @ z)
        assert abs(lp - lp_d) <= 1e-6 + 1e-6 * abs(lp_d), (lp, lp_d, order)


@pytest.mark.parametrize("ordering", ["forward", "reverse"])
def test_large_output_and_bottleneck_emissions(pkg, handle, ordering):
    """LargeOutputLGC (LGC:153-204, Dobs > Dlat) and BottleneckLGC (LGC:258-335) emissions: logpdf and the filtering distributions
    against the reference's OWN arithmetic for those types (oracle: posterior_and_lml_large / _bottleneck, incl. their jitters)."""
    rng = np.random.default_rng(31)
    from tests.util import random_psd, _stable
    T, D, M, K = 23, 3, 7, 2
    As = np.stack([_stable(0.9 * np.eye(D) + 0.1 * rng.standard_normal((D, D))) for _ in range(T)])
    as_ = rng.standard_normal((T, D)) * 0.3
    Qs = np.stack([random_psd(rng, D) for _ in range(T)])
    m0, P0 = rng.standard_normal(D), random_psd(rng, D, 0.5, 2.0)
    L = pkg.lgssm
    tr = L.GaussMarkovModel(ordering, As, as_, Qs, L.Gaussian(m0, P0))
    idx = range(T) if ordering == "forward" else range(T - 1, -1, -1)

    def run_oracle(update):
        m, P, lmls, ms = m0, P0, np.zeros(T), np.zeros((T, D))
        for t in idx:
            if ordering == "forward":
                m, P = O.predict(m, P, As[t], as_[t], Qs[t])
                m, P, lmls[t] = update(t, m, P)
                ms[t] = m
            else:
                m, P, lmls[t] = update(t, m, P)
                ms[t] = m
                m, P = O.predict(m, P, As[t], as_[t], Qs[t])
        return lmls.sum(), ms

    # LargeOutputLGC: Dobs = 7 > Dlat = 3
    Hs = rng.standard_normal((T, M, D)); hs = rng.standard_normal((T, M)) * 0.2
    Rs = np.stack([random_psd(rng, M, 0.1, 1.0) for _ in range(T)])
    y = rng.standard_normal((T, M))
    ref, ms_o = run_oracle(lambda t, m, P: O.posterior_and_lml_large(m, P, Hs[t], hs[t], Rs[t], y[t]))
    model = L.LGSSM(tr, L.LargeOutputEmissions(Hs, hs, Rs))
    lml = L.logpdf(model, y, handle)
    assert abs(lml - ref) <= 1e-6 * abs(ref), (lml, ref)
    ms, _ = L._filter(model, y, handle)
    np.testing.assert_allclose(ms, ms_o, rtol=1e-5, atol=1e-7)
    # BottleneckLGC: project D = 3 -> K = 2, fan out to M = 7
    Hp = rng.standard_normal((T, K, D)); hp = rng.standard_normal((T, K)) * 0.2
    Af = rng.standard_normal((T, M, K)); af = rng.standard_normal((T, M)) * 0.2
    ref, ms_o = run_oracle(lambda t, m, P: O.posterior_and_lml_bottleneck(m, P, Hp[t], hp[t], Af[t], af[t], Rs[t], y[t]))
    model = L.LGSSM(tr, L.BottleneckEmissions(Hp, hp, L.LargeOutputEmissions(Af, af, Rs)))
    lml = L.logpdf(model, y, handle)
    assert abs(lml - ref) <= 1e-6 * abs(ref), (lml, ref)
    ms, _ = L._filter(model, y, handle)
    np.testing.assert_allclose(ms, ms_o, rtol=1e-5, atol=1e-7)
    mu, var = L.marginals_diag(model, handle)        # predict_marginals(x, ::BottleneckLGC) (LGC:311-313)
    assert mu.shape == (T, M) and np.all(var > 0)
