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


@pytest.mark.parametrize("kernel", KERNELS, ids=KERNEL_IDS)
def test_sde_components_reproduce_the_per_step_construction(pkg, kernel):
    """gp.sde_components folds sums / scalings / time stretches into ONE (F, F0, H, P) — what tgp_lti_components takes. With SciPy's
    expm it must reproduce lgssm_components on an irregular grid: exp(F dt) for the later steps, exp(F0 * 1) for the first
    (lti_sde.jl:139, 361-373)."""
    from scipy.linalg import expm
    _, ko, kp = kernel
    k = kp(pkg)
    try:
        F, F0, H, P = pkg.gp.sde_components(k)
    except Exception:                      # kernels without a joint SDE form (products) keep the host construction
        pytest.skip("no joint SDE form")
    t = np.sort(np.random.default_rng(3).uniform(0.0, 7.0, 40))
    As, as_, Qs, Hs, hs, x0 = pkg.gp.lgssm_components(k, t)
    Psym = np.triu(P) + np.triu(P, 1).T
    dt = np.diff(t)
    for i in range(1, len(t)):
        A = expm(F * dt[i - 1])
        np.testing.assert_allclose(A, As[i], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(Psym - A @ Psym @ A.T, Qs[i], rtol=1e-9, atol=1e-12 * np.abs(P).max())
    np.testing.assert_allclose(expm(F0), As[0], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(H, pkg.gp._dense(Hs, len(t))[0])
    np.testing.assert_allclose(P, x0.P)


def test_device_steps_cross_the_boundary_by_address(pkg):
    """lgssm.DeviceSteps (what the device model builder returns) is marshalled as pointer + stride, nothing is copied."""
    import torch
    L = pkg.lgssm
    T, D = 5, 2
    A = torch.arange(T * D * D, dtype=torch.float64)
    Q = torch.ones(T * D * D, dtype=torch.float64)
    tr = L.GaussMarkovModel(L.Forward, L.DeviceSteps(A, T, (D, D)), L.Fill(np.zeros(D), T), L.DeviceSteps(Q, T, (D, D)),
                            L.Gaussian(np.zeros(D), np.eye(D)))
    m = L.LGSSM(tr, L.ScalarEmissions(L.Fill(np.array([1.0, 0.0]), T), L.Fill(np.zeros(()), T), L.Fill(np.array(0.1), T)))
    mm = L._Marshalled(m)
    assert mm.desc.A == A.data_ptr() and mm.desc.sA == D * D and mm.desc.Q == Q.data_ptr() and mm.desc.sQ == D * D
    assert mm.desc.sa == 0 and mm.desc.sH == 0
    # (T, D, D) in mathematical orientation: the device layout is column-major per step
    np.testing.assert_array_equal(L.DeviceSteps(A, T, (D, D)).numpy()[1], np.array([[4.0, 6.0], [5.0, 7.0]]))
    with pytest.raises(L.DimensionMismatch):
        L.DeviceSteps(A, T + 1, (D, D))


def test_shard_with_halo_layout(pkg):
    """The overlapped scatter: every rank's buffer is [3072 observations before the shard | shard], the view starts at the shard."""
    import torch
    from temporalgps_jl_b200 import sharded
    y = np.arange(40_000, dtype=np.float64)
    b = sharded.shard_bounds(len(y), 3)
    for r in range(3):
        buf, view = sharded.shard_with_halo(torch, y, b, r, "cpu")
        assert view.numel() == b[r + 1] - b[r] and view.data_ptr() - buf.data_ptr() == 8 * sharded.TGP_SHARD_HALO
        np.testing.assert_array_equal(view.numpy(), y[b[r]:b[r + 1]])
        if r > 0:
            np.testing.assert_array_equal(buf[:sharded.TGP_SHARD_HALO].numpy(), y[b[r] - sharded.TGP_SHARD_HALO:b[r]])
    with pytest.raises(ValueError):
        sharded.shard_with_halo(torch, y[:5000], sharded.shard_bounds(5000, 4), 1, "cpu")


def test_bottleneck_emissions_collapse(pkg):
    """BottleneckLGC (LGC:258-335) y | x ~ N(A (H x + h) + a, Q) is marshalled as the single conditional N((A H) x + (A h + a), Q)."""
    L = pkg.lgssm
    rng = np.random.default_rng(0)
    T, D, K, M = 4, 3, 2, 5
    H, h = rng.standard_normal((T, K, D)), rng.standard_normal((T, K))
    A, a = rng.standard_normal((T, M, K)), rng.standard_normal((T, M))
    Q = np.stack([np.eye(M) * (0.5 + i) for i in range(T)])
    em = L.BottleneckEmissions(H, h, L.LargeOutputEmissions(A, a, Q))
    sm = em.collapse(T)
    assert isinstance(sm, L.SmallOutputEmissions) and sm.M == M and sm.r_dense
    for t in range(T):
        np.testing.assert_allclose(sm.Hs[t], A[t] @ H[t])
        np.testing.assert_allclose(sm.hs[t], A[t] @ h[t] + a[t])
    fills = L.BottleneckEmissions(L.Fill(H[0], T), L.Fill(h[0], T), L.LargeOutputEmissions(L.Fill(A[0], T), L.Fill(a[0], T), L.Fill(Q[0], T)))
    sf = fills.collapse(T)
    assert isinstance(sf.Hs, L.Fill) and isinstance(sf.hs, L.Fill)       # time-invariant stays O(1)
    np.testing.assert_allclose(sf.Hs.value, A[0] @ H[0])


def test_value_and_gradient_is_fourth_order(pkg):
    f = lambda t: float(np.sin(t[0]) * np.exp(0.3 * t[1]) + t[0] * t[1] ** 2)     # noqa: E731
    th = np.array([0.7, -1.3])
    v, g = pkg.gp.value_and_gradient(f, th, rel_step=1e-2)
    ga = np.array([np.cos(th[0]) * np.exp(0.3 * th[1]) + th[1] ** 2, 0.3 * np.sin(th[0]) * np.exp(0.3 * th[1]) + 2 * th[0] * th[1]])
    assert v == f(th)
    np.testing.assert_allclose(g, ga, rtol=1e-8)


def test_halo_bounds_partition_the_series(pkg):
    from temporalgps_jl_b200 import sharded
    T, world, halo = 100_003, 7, 4096
    owned = []
    for r in range(world):
        lo, hi, lo_h, hi_h = sharded.halo_bounds(T, world, r, halo)
        assert 0 <= lo_h <= lo < hi <= hi_h <= T
        assert lo - lo_h == min(halo, lo) and hi_h - hi == min(halo, T - hi)
        owned.append((lo, hi))
    assert owned[0][0] == 0 and owned[-1][1] == T and all(a[1] == b[0] for a, b in zip(owned, owned[1:]))
    with pytest.raises(ValueError):        # wrong extended-shard length is refused before any device work
        sharded.posterior_marginals_sharded(pkg.gp, None, pkg.RegularSpacing(0.0, 0.01, T), 0.1, np.zeros(10), 1e-2, 1, world)


def test_scaling_and_squaring_taylor12_is_accurate_enough():
    """The matrix exponential k_lti_components uses (tgp_lti.cu): scale to |F dt| / 2^s <= 1/4 with s = ilogb(norm) + 3, degree-12 Taylor
    by Horner, s squarings — restated in NumPy and compared with SciPy's Pade-13 over twelve decades of dt for every base SDE (bound: 5e-14 max(1, |F dt|), i.e. the conditioning of the problem)."""
    from scipy.linalg import expm
    import math

    def expm_taylor(F, dt):
        X = F * dt
        nrm = np.abs(X).sum(axis=0).max()
        s = max(0, min(60, math.frexp(nrm)[1] - 1 + 3)) if nrm > 0.25 else 0
        X = F * math.ldexp(dt, -s)
        n = F.shape[0]
        E = X / 12.0 + np.eye(n)
        for k in range(11, 0, -1):
            E = (X @ E) / k + np.eye(n)
        for _ in range(s):
            E = E @ E
        return E

    lam3, lam5 = math.sqrt(3.0), math.sqrt(5.0)
    Fs = [np.array([[-1.0]]), np.array([[0.0, 1.0], [-lam3 ** 2, -2 * lam3]]),
          np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-lam5 ** 3, -3 * lam5 ** 2, -3 * lam5]]),
          np.array([[0.0, -2 * math.pi * 3], [2 * math.pi * 3, 0.0]])]          # a harmonic of the approximately periodic kernel
    for F in Fs:
        for dt in 10.0 ** np.arange(-9.0, 3.1, 0.5):
            A, B = expm_taylor(F, dt), expm(F * dt)
            nrm = np.abs(F * dt).sum(axis=0).max()          # squaring s ~ log2(norm) times loses ~ norm * eps (SciPy's Pade does too)
            assert np.abs(A - B).max() <= 5e-14 * max(1.0, nrm) * max(1.0, np.abs(B).max()), (F.shape, dt, np.abs(A - B).max())
