"""CPU checks of the C-ABI boundary: the library loads, exports every symbol include/tgp_b200.h declares, fails
loudly without a CUDA device (no CPU fallback), and the host mirror marshals arrays the way Julia lays them out."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "tgp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tgp_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(pkg):
    lib = C.CDLL(pkg._lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tgp_b200.h but not exported"
    assert set(pkg._lib.EXPORTS) == set(syms)
    assert b"sm_100a" in pkg._lib.lib().tgp_version()


def test_struct_layout_matches_header(pkg):
    d = pkg._lib.tgp_lgssm
    assert C.sizeof(d) == 4 + 4 + 8 + 4 + 4 + 6 * 16 + 16   # D M T ordering R_kind, 6 x (ptr, stride), m0 P0
    assert d.T.offset == 8 and d.A.offset == 24 and d.m0.offset == 24 + 96


def test_no_cpu_fallback_without_a_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.TGPError) as e:
        pkg.Handle(0)
    assert e.value.code == pkg._lib.TGP_ECUDA
    assert "no CPU path" in str(e.value)


def test_product_does_not_import_the_oracle():
    pk = os.path.join(ROOT, "temporalgps.jl_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_marshalling_is_column_major_with_fill_stride_zero(pkg):
    L = pkg.lgssm
    T, D = 5, 2
    A = np.arange(T * D * D, dtype=float).reshape(T, D, D)
    tr = L.GaussMarkovModel(L.Forward, A, L.Fill(np.zeros(D), T), L.Fill(np.eye(D), T), L.Gaussian(np.zeros(D), np.eye(D)))
    em = L.ScalarEmissions(L.Fill(np.array([1.0, 0.0]), T), L.Fill(np.zeros(()), T), np.full(T, 0.1))
    mm = L._Marshalled(L.LGSSM(tr, em))
    d = mm.desc
    assert (d.D, d.M, d.T, d.sA, d.sa, d.sQ, d.sH, d.sh, d.sR) == (2, 1, 5, 4, 0, 0, 0, 0, 1)
    flat = np.ctypeslib.as_array(C.cast(d.A, C.POINTER(C.c_double)), shape=(T * D * D,))
    # A[t][i, j] is stored at t*4 + i + 2*j (Julia SMatrix order)
    assert flat[1 * 4 + 0 + 2 * 1] == A[1][0, 1] and flat[1 * 4 + 1 + 2 * 0] == A[1][1, 0]


def test_dimension_mismatch_raises_before_any_device_work(pkg):
    L = pkg.lgssm
    fx = pkg.to_sde(pkg.GP(pkg.Matern32Kernel()))(pkg.RegularSpacing(0.0, 0.1, 10), 0.1)
    with pytest.raises(pkg.DimensionMismatch):
        L.logpdf(fx.build_lgssm(), np.zeros(9))


def test_missing_transform_host_side(pkg):
    """missings.jl:25-53: y := 0, R := 1e15 at masked entries; dispatch on the masked-array TYPE."""
    L = pkg.lgssm
    fx = pkg.to_sde(pkg.GP(pkg.Matern32Kernel()))(pkg.RegularSpacing(0.0, 0.1, 6), 0.1)
    y = np.ma.masked_array(np.arange(6.0), mask=[0, 1, 0, 0, 1, 0])
    m2, y2, n = L.transform_model_and_obs(fx.build_lgssm(), y)
    assert n == 2 and list(y2) == [0, 0, 2, 3, 0, 5]
    assert list(m2.emissions.Rs) == [0.1, 1e15, 0.1, 0.1, 1e15, 0.1]
    m3, y3, n3 = L._maybe_missing(fx.build_lgssm(), np.arange(6.0))
    assert n3 == 0 and isinstance(m3.emissions.Rs, L.Fill)


def test_merge_datasets_matches_reference_semantics(pkg):
    """posterior_lti_sde.jl:97-123: stable sort, train/predict index maps."""
    x, S, ys, tr, pr = pkg.gp.merge_datasets(np.array([0.0, 1.0, 2.0]), np.array([0.5, 2.5]), pkg.Fill(0.1, 3), pkg.Fill(1e15, 2),
                                            np.array([1.0, 2.0, 3.0]), np.full(2, np.nan))
    assert list(x) == [0.0, 0.5, 1.0, 2.0, 2.5]
    assert list(tr) == [0, 2, 3] and list(pr) == [1, 4]
    assert list(np.ma.getmaskarray(ys)) == [False, True, False, False, True]
    assert list(S) == [0.1, 1e15, 0.1, 0.1, 1e15]
