"""The C ABI driven from plain C (tests/abi_c/abi_layout.c) with buffers laid out as the Julia glue lays them out (`Fill` -> stride 0,
`reinterpret(Float64, ::Vector{SMatrix})`, `Vector{Gaussian}` records written in place): the stand-in for executing the glue."""
import os
import subprocess

import numpy as np
import pytest

from oracle import tgp_oracle as O
from tests.util import random_lgssm, sample_y

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "abi_c", "abi_layout.c")
EXE = os.path.join(ROOT, "tests", "abi_c", "_build", "abi_layout")
LIBDIR = os.path.join(ROOT, "temporalgps.jl_b200")


def build_exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-std=c11", "-o", EXE, SRC, "-L" + LIBDIR, "-ltgpb200", "-lm", "-Wl,-rpath," + LIBDIR], check=True)
    return EXE


def test_abi_c_program_compiles_and_links(pkg):
    """(CPU) every symbol the C program uses resolves against the built library."""
    assert os.path.exists(build_exe())


def _write(path, m, y, tv):
    ms, Ps, lmls = O.filter_(m, y)
    T, D = m.T, m.D
    n = T if tv else 1
    cm = lambda M3: np.ascontiguousarray(np.swapaxes(np.asarray(M3)[:n], 1, 2))     # column-major blocks  # noqa: E731
    recs = np.concatenate([ms, np.swapaxes(Ps, 1, 2).reshape(T, D * D)], axis=1)
    parts = [np.array([D, T, 1.0 if tv else 0.0]), cm(m.As), np.asarray(m.as_)[:n], cm(m.Qs), np.asarray(m.Hs)[:n], np.asarray(m.hs)[:n],
             np.asarray(m.Rs)[:n], np.asarray(m.m0), np.asarray(m.P0).T, y, np.array([lmls.sum()]), recs]
    with open(path, "wb") as fh:
        for p in parts:
            fh.write(np.ascontiguousarray(p, dtype=np.float64).tobytes())


@pytest.mark.gpu
@pytest.mark.parametrize("D,T,tv", [(3, 1000, True), (2, 49, True), (3, 100_000, False), (4, 5000, False)])
def test_abi_from_c_with_julia_layouts(pkg, tmp_path, D, T, tv):
    rng = np.random.default_rng(D * 7 + T)
    if tv:
        m = random_lgssm(rng, T, D, "forward", True)
    else:
        k = {2: O.Matern32(), 3: O.Matern52(), 4: O.Sum([O.Matern32(), O.Matern32()])}[D]
        m = O.build_lgssm(k, O.RegularSpacing(0.0, 0.05, T), 0.2)
    y = sample_y(rng, m) if tv else O.sample_prior(m, rng)
    path = str(tmp_path / "case.bin")
    _write(path, m, y, tv)
    p = subprocess.run([build_exe(), path], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "ok" in p.stdout
