"""GPU parity tests of the one-launch steady-state logpdf (tgp_fir.cuh, `k_fir_logpdf`) through the C ABI, against the sequential
C oracle on the same seeded inputs. Tolerance: north_star's 1e-6 relative on logpdf (measured ~1e-14)."""
import numpy as np
import pytest

from oracle import c_oracle, tgp_oracle as O

pytestmark = pytest.mark.gpu

LML_RTOL = 1e-6
TIGHT = 1e-11   # what the path actually delivers; a regression here is a bug even if 1e-6 still holds


def _models(pkg):
    return {
        "m12": (pkg.Matern12Kernel(), O.Matern12(), 0.05),
        "m32": (pkg.Matern32Kernel(), O.Matern32(), 0.1),
        "m52": (pkg.Matern52Kernel(), O.Matern52(), 0.01),
        "sum_m32_m12": (pkg.Matern32Kernel() + 0.5 * pkg.Matern12Kernel(), O.Sum([O.Matern32(), O.Scaled(0.5, O.Matern12())]), 0.02),
        "sum_m52_m12": (pkg.Matern52Kernel() + 0.3 * pkg.Matern12Kernel(), O.Sum([O.Matern52(), O.Scaled(0.3, O.Matern12())]), 0.02),
    }


def _launches(handle, fn):
    c0 = handle.counters()["launches"]
    out = fn()
    return out, handle.counters()["launches"] - c0


@pytest.mark.parametrize("name", ["m12", "m32", "m52", "sum_m32_m12", "sum_m52_m12"])
@pytest.mark.parametrize("T", [4096, 5000, 20011, 65536, 300_001])
def test_fir_logpdf_matches_oracle(pkg, handle, name, T):
    kp, ko, dt = _models(pkg)[name]
    mo = O.build_lgssm(ko, O.RegularSpacing(0.0, dt, T), 0.1)
    y = O.sample_prior(mo, np.random.default_rng(T))
    ref = c_oracle.logpdf(c_oracle.Model.from_lgssm(mo), y)
    fx = pkg.to_sde(pkg.GP(kp))(pkg.RegularSpacing(0.0, dt, T), 0.1)
    lml, n = _launches(handle, lambda: pkg.gp.logpdf(fx, y))
    assert abs(lml - ref) <= TIGHT * abs(ref), (lml, ref)
    assert n == 1, f"{n} launches: the call did not take the one-launch path"
    # same answer as the general scan (the algorithm north_star names)
    handle.set_algo(pkg.TGP_ALGO_SCAN)
    try:
        lml_scan = pkg.gp.logpdf(fx, y)
    finally:
        handle.set_algo(pkg.TGP_ALGO_AUTO)
    assert abs(lml - lml_scan) <= LML_RTOL * abs(ref)


@pytest.mark.parametrize("off", [0, 1, 2, 3])
def test_fir_any_alignment_device_resident(pkg, handle, off):
    """y anywhere in HBM (8-byte aligned only): the plan starts the steady phase on a 32-byte boundary."""
    import torch
    T = 150_000
    mo = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.01, T), 0.1)
    y = O.sample_prior(mo, np.random.default_rng(off))
    ref = c_oracle.logpdf(c_oracle.Model.from_lgssm(mo), y)
    buf = torch.zeros(T + 8, dtype=torch.float64, device="cuda")
    buf[off:off + T] = torch.from_numpy(y).cuda()
    mm = pkg.lgssm._Marshalled(pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 0.01, T), 0.1).build_lgssm())
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    handle.logpdf(mm.desc, buf[off:off + T], out)        # device destination: enqueued only
    handle.synchronize()
    lml = float(out.item())
    assert abs(lml - ref) <= TIGHT * abs(ref), (lml, ref)


def test_fir_repeated_calls_plan_cache_and_reproducibility(pkg, handle):
    """Consecutive calls reuse the workspace (epoch-tagged tile words); a new model or length rebuilds the plan; the result is
    bit-reproducible (static tile assignment, fixed-order reductions)."""
    rng = np.random.default_rng(11)
    seen = {}
    for rep in range(3):
        for (dt, s2, T) in [(0.01, 0.1, 100_000), (0.02, 0.3, 100_000), (0.01, 0.1, 40_000), (0.01, 0.1, 100_000)]:
            mo = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, dt, T), s2)
            y = O.sample_prior(mo, np.random.default_rng(int(1000 * dt) + T))
            ref = c_oracle.logpdf(c_oracle.Model.from_lgssm(mo), y)
            fx = pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, dt, T), s2)
            lml = pkg.gp.logpdf(fx, y)
            assert abs(lml - ref) <= TIGHT * abs(ref)
            key = (dt, s2, T)
            assert seen.setdefault(key, lml) == lml, "not bit-reproducible"


def test_fir_full_size_config2_properties(pkg, handle):
    """BASELINE config 2 at full size (T = 1e7): agreement with the sequential oracle and a size-independent property —
    log p(y) of the series = log p(first half) + log p(second half | first half), the second term obtained as the difference of two
    runs of different length."""
    T = 10_000_000
    mo = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.01, T), 0.1)
    rng = np.random.default_rng(20261017 + 2)
    y = np.cumsum(rng.standard_normal(T)) * 0.01
    y = y - np.linspace(0, y[-1], T) + 0.3 * rng.standard_normal(T)
    ref = c_oracle.logpdf(c_oracle.Model.from_lgssm(mo), y)
    f = pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))
    lml, n = _launches(handle, lambda: pkg.gp.logpdf(f(pkg.RegularSpacing(0.0, 0.01, T), 0.1), y))
    assert n == 1
    assert abs(lml - ref) <= TIGHT * abs(ref), (lml, ref)
    half = T // 2
    lml_half = pkg.gp.logpdf(f(pkg.RegularSpacing(0.0, 0.01, half), 0.1), y[:half])
    ref_half = c_oracle.logpdf(c_oracle.Model.from_lgssm(O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.01, half), 0.1)), y[:half])
    assert abs((lml - lml_half) - (ref - ref_half)) <= 1e-9 * abs(ref)


def test_fir_not_positive_definite_reports_the_step(pkg, handle):
    T = 8192
    fx = pkg.to_sde(pkg.GP(pkg.Matern32Kernel()))(pkg.RegularSpacing(0.0, 0.1, T), 0.1)
    model = fx.build_lgssm()
    model.emissions.Rs = pkg.lgssm.Fill(np.array(-10.0), T)
    with pytest.raises(pkg.PosDefException) as ei:
        pkg.lgssm.logpdf(model, np.zeros(T), handle)
    assert "time index 0" in str(ei.value)


def test_fir_slow_forgetting_falls_back(pkg, handle):
    """A grid so fine that the filter's memory exceeds the 3-tile look-back: the plan declines, the two-phase / general kernels run."""
    T = 70_000
    mo = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 1e-4, T), 0.1)
    y = O.sample_prior(mo, np.random.default_rng(3))
    ref = c_oracle.logpdf(c_oracle.Model.from_lgssm(mo), y)
    fx = pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 1e-4, T), 0.1)
    lml, n = _launches(handle, lambda: pkg.gp.logpdf(fx, y))
    assert n > 1
    assert abs(lml - ref) <= LML_RTOL * abs(ref)
