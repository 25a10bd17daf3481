#!/usr/bin/env python
"""Probe of the tensor-core path on the GPU box: accuracy of the 3xTF32 contraction kernel per shape, then accuracy and
time per step of the FP32-storage dense step against the FP64 dense step and the oracle (config-5 shapes)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
from oracle import tgp_oracle as O

pkg = g.load_package()
h = pkg.default_handle(0)
for K, Mx, N in [(32, 128, 64), (128, 128, 64), (256, 256, 768), (768, 768, 768), (768, 768, 256), (100, 70, 50), (33, 129, 65)]:
    rng = np.random.default_rng(K + Mx + N)
    X = rng.standard_normal((K, Mx)).astype(np.float32)
    Y = rng.standard_normal((K, N)).astype(np.float32)
    C = h.tc_gemm(X, Y)
    ref = X.astype(np.float64).T @ Y.astype(np.float64)
    ref32 = X.T @ Y
    Xp = np.abs(X); Yp = np.abs(Y)
    Cp = h.tc_gemm(Xp, Yp)
    refp = Xp.astype(np.float64).T @ Yp.astype(np.float64)
    print(json.dumps({"gemm": [K, Mx, N], "max_abs_err_over_sqrtK": float(np.max(np.abs(C - ref)) / np.sqrt(K)),
                      "numpy_f32_err_over_sqrtK": float(np.max(np.abs(ref32 - ref)) / np.sqrt(K)),
                      "positive_operands_max_rel_err": float(np.max(np.abs(Cp - refp) / refp)),
                      "numpy_f32_positive_rel_err": float(np.max(np.abs(Xp.T @ Yp - refp) / refp))}), flush=True)


def run(Nr, T, Tcpu):
    r = np.linspace(-3.0, 3.0, Nr)
    rng = np.random.default_rng(20261017 + 5)
    mo = O.build_lgssm_separable(O.SqExp(), O.Matern52(), r, O.RegularSpacing(0.0, 0.01, Tcpu), 0.1)
    y = rng.standard_normal((T, Nr))
    y[:Tcpu] = O.sample_prior(mo, rng)
    ref = O.logpdf_steps(mo, y[:Tcpu])
    out = {"Nr": Nr, "D": 3 * Nr, "M": Nr, "T": T}
    for name, dt in (("f64", np.float64), ("tf32x3", np.float32)):
        fx = pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern52Kernel())), pkg.ArrayStorage(dt))(
            pkg.RectilinearGrid(r, pkg.RegularSpacing(0.0, 0.01, T)), 0.1)
        model = fx.build_lgssm()
        hh = fx._handle()
        lml, steps = pkg.lgssm.logpdf(model, y, hh, per_step=True)
        t0 = time.perf_counter()
        lml = pkg.lgssm.logpdf(model, y, hh)
        dtm = time.perf_counter() - t0
        out[name] = {"ms_per_step": dtm / T * 1e3, "lml_steps_max_rel_err_vs_oracle": float(np.max(np.abs(steps[:Tcpu] - ref) / np.abs(ref))),
                     "lml_sum_rel_err": float(abs(steps[:Tcpu].sum() - ref.sum()) / abs(ref.sum()))}
        fx16 = pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern52Kernel())), pkg.ArrayStorage(dt))(
            pkg.RectilinearGrid(r, pkg.RegularSpacing(0.0, 0.01, 16)), 0.1)
        hh.set_timing(True)
        pkg.lgssm.logpdf(fx16.build_lgssm(), y[:16], hh)
        out[name]["kernels_us_per_step"] = {n: round(ms / 16 * 1e3, 2) for n, ms, c in hh.timing()}
        hh.set_timing(False)
    h.set_dense_math(0)
    print(json.dumps(out), flush=True)


import sys as _s
if '--big-only' not in _s.argv:
    run(24, 200, 200)
    run(64, 300, 100)
run(256, 200, 12)
