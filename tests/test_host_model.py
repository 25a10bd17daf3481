"""The host-side GP -> LGSSM construction of the package (gp.py, the CALLER of the hot path) against the oracle's
independent restatement of src/gp/lti_sde.jl, for the reference's whole kernel list. CPU only."""
import numpy as np
import pytest

from oracle import tgp_oracle as O
from tests.kernels import KERNEL_IDS, KERNELS, MEANS, inputs


@pytest.mark.parametrize("regular", [True, False], ids=["regular", "irregular"])
@pytest.mark.parametrize("kernel", KERNELS, ids=KERNEL_IDS)
def test_lgssm_components_match_oracle(pkg, kernel, regular):
    _, ko, kp = kernel
    tp, to = inputs(regular, pkg), inputs(regular, O)
    a = pkg.gp.lgssm_components(kp(pkg), tp)
    b = O.lgssm_components(ko(), to)
    T = len(tp)
    for x, y in zip(a[:5], b[:5]):
        np.testing.assert_allclose(pkg.gp._dense(x, T), np.asarray(y), rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(a[5].m, b[5][0])
    np.testing.assert_allclose(a[5].P, b[5][1], rtol=1e-14)
    if regular:   # Fill structure survives (O(1) memory, lti_sde.jl:148-160)
        assert isinstance(a[0], pkg.Fill) and isinstance(a[2], pkg.Fill)


@pytest.mark.parametrize("mean", MEANS, ids=[m[0] for m in MEANS])
def test_build_lgssm_with_means_and_noise(pkg, mean):
    rng = np.random.default_rng(0)
    s2 = rng.uniform(size=13) + 0.1
    fx = pkg.to_sde(pkg.GP(pkg.Matern52Kernel(), mean[1]))(pkg.RegularSpacing(0.0, 0.3, 13), s2)
    m = fx.build_lgssm()
    mo = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.3, 13), s2, mean[1])
    np.testing.assert_allclose(pkg.gp._dense(m.emissions.hs, 13), mo.hs)
    np.testing.assert_allclose(pkg.gp._dense(m.emissions.Rs, 13), mo.Rs)
    assert pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(np.arange(3.0)).noise == 1e-12   # lti_sde.jl:27-29
