#!/usr/bin/env python
"""Generates tests/golden/*.npz — small input/output vectors of the hot path produced by the CPU oracle
(oracle/tgp_oracle.py, cross-checked against the C restatement). The reference itself is Julia and cannot be run
here, so these are ORACLE outputs, pinned by tests/test_oracle_pins.py to the reference's own
equivalence tests (dense GP). Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle, tgp_oracle as O  # noqa: E402
from tests.util import random_lgssm, sample_y  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, model, y, Rn):
    ms, Ps, lmls = O.filter_(model, y)
    d = dict(ordering=model.ordering, As=np.array(model.As), as_=np.array(model.as_), Qs=np.array(model.Qs), m0=model.m0, P0=model.P0,
             Hs=np.array(model.Hs), hs=np.array(model.hs), Rs=np.array(model.Rs), y=y, ti=np.array(model.As.strides[0] == 0),
             lml=O.logpdf(model, y), lml_steps=lmls, m_f=ms, P_f=Ps)
    mu, var = O.marginals(model)
    d.update(prior_mean=mu, prior_var=var)
    if model.ordering == "forward":
        post = O.posterior(model, y)
        pm, pv = O.marginals(O.replace_observation_noise_cov(post, Rn))
        d.update(R_new=Rn, G=post.As, g=post.as_, Sig=post.Qs, post_mean=pm, post_var=pv)
        c = c_oracle.filter(c_oracle.Model.from_lgssm(model), y)   # second restatement must agree before we pin
        assert np.allclose(c["lml_steps"], lmls, rtol=1e-11, atol=1e-12)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def main():
    rng = np.random.default_rng(20261017)
    # (1) README example shape: Matern32, RegularSpacing(0, 0.1, .), sigma^2 = 0.1 (config 1, shortened)
    T = 500
    m = O.build_lgssm(O.Matern32(), O.RegularSpacing(0.0, 0.1, T), 0.1)
    save("cfg1_matern32_T500", m, O.sample_prior(m, rng), np.full(T, 1e-2))
    # (2) config-2 model, short
    T = 300
    m = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.01, T), 0.1)
    save("cfg2_matern52_T300", m, O.sample_prior(m, rng), np.full(T, 1e-2))
    # (3) irregular inputs, heteroscedastic noise, custom mean (time-varying everything), D = 3
    T = 64
    t = np.sort(rng.uniform(0, 10, T))
    m = O.build_lgssm(1.3 * O.Matern52().stretch(0.7), t, rng.uniform(0.05, 0.5, T), lambda x: 2.0 * x)
    save("tv_matern52_irregular_T64", m, O.sample_prior(m, rng), rng.uniform(0.01, 0.3, T))
    # (4) random LGSSMs in the style of the reference's fixtures, both orderings
    for D in (1, 2, 4):
        for ordering in ("forward", "reverse"):
            m = random_lgssm(rng, 49, D, ordering, True)
            save(f"random_D{D}_{ordering}_T49", m, sample_y(rng, m), rng.uniform(0.01, 0.3, 49))
    # (5) sum kernel D = 5 (Matern32 + Matern52), regular
    T = 200
    m = O.build_lgssm(O.Matern32() + 0.5 * O.Matern52().stretch(2.0), O.RegularSpacing(0.0, 0.05, T), 0.2)
    save("sum_m32_m52_T200", m, O.sample_prior(m, rng), np.full(T, 0.05))


if __name__ == "__main__":
    main()
