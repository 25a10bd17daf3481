"""CPU checks of the oracle's vector-observation additions: marginals_diag (lgssm.jl:125-141) and element-wise missing observations
(linear_gaussian_conditionals.jl:143-151), pinned by the identities the reference's own tests use."""
import numpy as np
import pytest

from oracle import tgp_oracle as O
from tests.test_gpu_parity import _random_vector_lgssm


@pytest.mark.parametrize("ordering", ["forward", "reverse"])
def test_marginals_diag_is_the_diagonal(ordering):
    m = _random_vector_lgssm(np.random.default_rng(1), 17, 4, 3, ordering, True)
    mu, cov = O.marginals(m)
    mu_d, var_d = O.marginals_diag(m)
    np.testing.assert_array_equal(mu, mu_d)
    np.testing.assert_array_equal(var_d, np.einsum("tii->ti", cov))


@pytest.mark.parametrize("ordering", ["forward", "reverse"])
def test_missing_entries_equal_dropping_them(ordering):
    """An entry with variance 1e15 carries (numerically) no information: the filter equals the one of the model whose emission at
    that step simply lacks the entry, and the compensated logpdf equals that model's (test/models/missings.jl:94-115 in spirit)."""
    rng = np.random.default_rng(2)
    T, D, M = 15, 3, 3
    m = _random_vector_lgssm(rng, T, D, M, ordering, False)
    y = O.sample_prior(O.LGSSM("forward", m.As, m.as_, m.Qs, m.m0, m.P0, m.Hs, m.hs, m.Rs), rng)
    y_nan = y.copy()
    drop = rng.random((T, M)) < 0.3
    drop[:, 0] = False                     # keep one entry per row so every step still observes something
    y_nan[drop] = np.nan
    lml = O.logpdf_missing(m, y_nan)
    # reference recursion with the missing entries removed from (H, h, R, y) step by step
    mm, P = m.m0, m.P0
    tot = 0.0
    for t in m.indices():
        keep = ~drop[t]
        H, h, R, yt = m.Hs[t][keep], m.hs[t][keep], m.Rs[t][np.ix_(keep, keep)], y[t][keep]
        if ordering == "forward":
            mm, P = O.predict(mm, P, m.As[t], m.as_[t], m.Qs[t])
            mm, P, l = O.posterior_and_lml_small(mm, P, H, h, R, yt)
        else:
            mm, P, l = O.posterior_and_lml_small(mm, P, H, h, R, yt)
            mm, P = O.predict(mm, P, m.As[t], m.as_[t], m.Qs[t])
        tot += l
    assert abs(lml - tot) <= 1e-6 * abs(tot)
    md = _random_vector_lgssm(rng, T, D, M, ordering, True)
    with pytest.raises(TypeError):
        O.logpdf_missing(md, y_nan)
