"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle):
CPU: both oracle restatements still reproduce them; GPU: the CUDA path matches them through the C ABI."""
import glob
import os

import numpy as np
import pytest

from oracle import c_oracle, tgp_oracle as O
from tests.util import to_pkg_model

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
IDS = [os.path.basename(f)[:-4] for f in FILES]


def _load(path):
    g = np.load(path, allow_pickle=False)
    T = len(g["y"])
    ti = bool(g["ti"])

    def per_step(x):
        return np.broadcast_to(x[0], x.shape) if ti else x

    m = O.LGSSM(str(g["ordering"]), per_step(g["As"]), per_step(g["as_"]), per_step(g["Qs"]), g["m0"], g["P0"], per_step(g["Hs"]),
                per_step(g["hs"]), per_step(g["Rs"]))
    return g, m


def test_golden_files_exist():
    assert len(FILES) >= 10


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_oracles_reproduce_golden(path):
    g, m = _load(path)
    ms, Ps, lmls = O.filter_(m, g["y"])
    np.testing.assert_allclose(lmls, g["lml_steps"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(ms, g["m_f"], rtol=1e-12, atol=1e-13)
    r = c_oracle.filter(c_oracle.Model.from_lgssm(m), g["y"])
    np.testing.assert_allclose(r["lml_steps"], g["lml_steps"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(r["P"], g["P_f"], rtol=1e-10, atol=1e-12)
    assert abs(r["lml"] - float(g["lml"])) <= 1e-12 * abs(float(g["lml"]))


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_cuda_path_matches_golden(pkg, handle, path):
    g, m = _load(path)
    pm = to_pkg_model(pkg, m)
    y = np.array(g["y"])
    lml, steps = pkg.lgssm.logpdf(pm, y, handle, per_step=True)
    assert abs(lml - float(g["lml"])) <= 1e-6 * abs(float(g["lml"]))
    np.testing.assert_allclose(steps, g["lml_steps"], rtol=1e-6, atol=1e-8)
    ms, Ps = pkg.lgssm._filter(pm, y, handle)
    np.testing.assert_allclose(ms, g["m_f"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(Ps, g["P_f"], rtol=1e-5, atol=1e-10)
    mu, var = pkg.lgssm.marginals(pm, handle)
    np.testing.assert_allclose(mu, g["prior_mean"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(var, g["prior_var"], rtol=1e-9)
    if str(g["ordering"]) == "forward":
        mu, var = pkg.lgssm.posterior_marginals(pm, y, np.array(g["R_new"]), handle)
        np.testing.assert_allclose(mu, g["post_mean"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(var, g["post_var"], rtol=1e-5)
        post = pkg.lgssm.posterior(pm, y, handle)
        np.testing.assert_allclose(post.transitions.As, g["G"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(post.transitions.Qs, g["Sig"], rtol=1e-5, atol=1e-9)
