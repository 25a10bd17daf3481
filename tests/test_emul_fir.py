"""CPU test of the one-launch steady-state logpdf (tgp_fir.cuh): tests/emul/emul_fir.cpp compiles the kernel's plan builder and
per-lane arithmetic (tgp_fir_plan.h) with g++ and drives them with the kernel's tile / look-back structure. Compared with the
sequential oracle. (Test harness only.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import c_oracle, tgp_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul_fir.cpp")
SO = os.path.join(HERE, "emul", "_build", "libemul_fir.so")
CSRC = os.path.join(HERE, "..", "temporalgps.jl_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    deps = [SRC, os.path.join(CSRC, "tgp_fir_plan.h"), os.path.join(CSRC, "tgp_math.cuh")]
    if not os.path.exists(SO) or max(map(os.path.getmtime, deps)) > os.path.getmtime(SO):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO], check=True)
    L = C.CDLL(SO)
    L.emul_fir_logpdf.restype = C.c_int
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def run(emul, mo, y, align=0, first=True, halo=None, tol=1e-13):
    D = mo.D
    A = np.ascontiguousarray(np.asarray(mo.As[0]).T)       # column-major
    a = np.ascontiguousarray(mo.as_[0]); Q = np.ascontiguousarray(np.asarray(mo.Qs[0]).T)
    H = np.ascontiguousarray(mo.Hs[0]); m0 = np.ascontiguousarray(mo.m0); P0 = np.ascontiguousarray(np.asarray(mo.P0).T)
    lml = np.zeros(1); info = np.zeros(8, dtype=np.int64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    rc = emul.emul_fir_logpdf(D, _p(A), _p(a), _p(Q), _p(H), C.c_double(float(mo.hs[0])), C.c_double(float(mo.Rs[0])), _p(m0), _p(P0),
                              C.c_int64(len(y)), _p(y), C.c_double(tol), align, 1 if first else 0,
                              _p(np.ascontiguousarray(halo)) if halo is not None else None, _p(lml), _p(info))
    return rc, float(lml[0]), info


KERNELS = {
    "matern12": (O.Matern12(), 0.05), "matern32": (O.Matern32(), 0.1), "matern52": (O.Matern52(), 0.01),
    "sum32+12": (O.Sum([O.Matern32(), O.Scaled(0.5, O.Matern12())]), 0.02),
}


@pytest.mark.parametrize("name", list(KERNELS))
@pytest.mark.parametrize("T,align", [(4096, 0), (5000, 1), (20011, 3), (65536, 2)])
def test_fir_emulation_matches_sequential_filter(emul, name, T, align):
    k, dt = KERNELS[name]
    mo = O.build_lgssm(k, O.RegularSpacing(0.0, dt, T), 0.1)
    y = O.sample_prior(mo, np.random.default_rng(T + align))
    rc, lml, info = run(emul, mo, y, align)
    assert rc == 0, info
    ref = c_oracle.logpdf(c_oracle.Model.from_lgssm(mo), y)
    assert abs(lml - ref) <= 1e-11 * abs(ref), (lml, ref, info)
    assert ((align + info[1]) & 3) == 0          # the first steady step is 32-byte aligned


def test_fir_emulation_time_shard_from_halo(emul):
    """A shard with rank > 0: the state entering it comes from pass A over the nb tiles before it."""
    T, cut = 60000, 30000
    mo = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.01, T), 0.1)
    y = O.sample_prior(mo, np.random.default_rng(5))
    cm = c_oracle.Model.from_lgssm(mo)
    ref_all = c_oracle.logpdf(cm, y)
    mo1 = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.01, cut), 0.1)
    ref_first = c_oracle.logpdf(c_oracle.Model.from_lgssm(mo1), y[:cut])
    mo2 = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.01, T - cut), 0.1)
    rc, _, info = run(emul, mo2, y[cut:], 0, True)
    nb = int(info[2])
    rc, lml2, info = run(emul, mo2, y[cut:], 0, False, halo=y[cut - nb * 1024:cut])
    assert rc == 0 and info[1] == 0
    assert abs((ref_first + lml2) - ref_all) <= 1e-11 * abs(ref_all), (ref_first + lml2, ref_all)


def test_fir_plan_rejects_what_it_cannot_do(emul):
    # a grid so fine that the filter forgets too slowly for a 3-tile look-back -> status 1 (the two-phase kernel takes it)
    mo = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 1e-4, 8192), 0.1)
    y = np.zeros(8192)
    rc, _, info = run(emul, mo, y)
    assert rc == 1
    # negative noise variance -> not positive definite at step 0 (the reference's cholesky throws there)
    mo = O.build_lgssm(O.Matern32(), O.RegularSpacing(0.0, 0.1, 8192), 0.1)
    mo.Rs = np.broadcast_to(np.array([-10.0]), (8192,))
    rc, _, info = run(emul, mo, y)
    assert rc == 2 and info[4] == 0
