#!/usr/bin/env python
"""BASELINE config 5 at full size: Separable(SE, Matern52) on 256 spatial points x T = 100 000 regular times (D = 768, M = 256),
logpdf. FP64 (library GEMMs) and FP32 storage (tcgen05 3xTF32), steady-state switch on and off (off on a shorter series),
CPU oracle per-step time beside them.   python tools/cfg5_bench.py [--T 100000] [--Nr 256]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g
from oracle import tgp_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=100_000)
ap.add_argument("--Nr", type=int, default=256)
ap.add_argument("--Tfull", type=int, default=300, help="series length for the runs with the steady-state switch disabled")
ap.add_argument("--Tcpu", type=int, default=6)
a = ap.parse_args()
pkg = g.load_package()
Nr, T = a.Nr, a.T
D, M = 3 * Nr, Nr
r = np.linspace(-3.0, 3.0, Nr)
rng = np.random.default_rng(20261017 + 5)
y = rng.standard_normal((T, Nr))
mo = O.build_lgssm_separable(O.SqExp(), O.Matern52(), r, O.RegularSpacing(0.0, 0.01, a.Tcpu), 0.1)
t0 = time.perf_counter()
ref = O.logpdf_steps(mo, y[:a.Tcpu])
t_cpu = (time.perf_counter() - t0) / a.Tcpu
flops = 4.0 * D ** 3 + 2.0 * M * D * D + 2.0 * M * M * D + M ** 3 / 3.0 + 1.0 * M * M * D + 2.0 * D * D * M
out = {"config": f"cfg5 Separable(SE, Matern52) Nr={Nr} (D={D}, M={M}) logpdf", "T": T, "cpu_oracle_ms_per_step": t_cpu * 1e3,
       "cpu_sample_steps": a.Tcpu, "dense_flop_per_step": flops}
yd = torch.from_numpy(y).cuda()
for name, dt in (("f64", np.float64), ("f32_tf32x3", np.float32)):
    def fxT(n):
        return pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern52Kernel())), pkg.ArrayStorage(dt))(
            pkg.RectilinearGrid(r, pkg.RegularSpacing(0.0, 0.01, n)), 0.1)
    res = {}
    fx = fxT(T)
    hh = fx._handle()
    mm = pkg.lgssm._Marshalled(fx.build_lgssm())
    lml = np.zeros(1)
    c0 = hh.counters()
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hh.logpdf(mm.desc, yd, lml)
        dtm = time.perf_counter() - t0
    c1 = hh.counters()
    res["steady_switch"] = {"s_per_logpdf": dtm, "us_per_step": dtm / T * 1e6, "steps_per_s": T / dtm, "lml": float(lml[0]),
                            "launches_per_logpdf": (c1["launches"] - c0["launches"]) / 2}
    # switch off: every step runs the full covariance update (what a time-varying model costs)
    fx2 = fxT(a.Tfull)
    mm2 = pkg.lgssm._Marshalled(fx2.build_lgssm())
    hh.set_algo(pkg.TGP_ALGO_SCAN)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hh.logpdf(mm2.desc, yd[:a.Tfull], lml)
        dt2 = time.perf_counter() - t0
    hh.set_algo(pkg.TGP_ALGO_AUTO)
    res["full_steps"] = {"T": a.Tfull, "us_per_step": dt2 / a.Tfull * 1e6, "dense_equiv_tflops": flops * a.Tfull / dt2 / 1e12}
    # parity of the first steps against the oracle
    fx3 = fxT(a.Tcpu)
    l3, s3 = pkg.lgssm.logpdf(fx3.build_lgssm(), y[:a.Tcpu], hh, per_step=True)
    res["lml_step_max_rel_err_vs_oracle"] = float(np.max(np.abs(s3 - ref) / np.abs(ref)))
    out[name] = res
    hh.set_dense_math(0)
out["f32_vs_f64_lml_rel_diff"] = abs(out["f32_tf32x3"]["steady_switch"]["lml"] - out["f64"]["steady_switch"]["lml"]) / abs(out["f64"]["steady_switch"]["lml"])
print(json.dumps(out))
