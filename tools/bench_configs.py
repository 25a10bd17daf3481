#!/usr/bin/env python
"""Timings of the secondary BASELINE configs (1 and 3) on one GPU, with the CPU oracle beside them and the parity
error of the timed call. Prints one JSON object per config.   python tools/bench_configs.py [--T3 1000000]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g
from oracle import c_oracle, tgp_oracle as O

pkg = g.load_package()
h = pkg.default_handle(0)


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, r


def cfg1():
    T = 10_000
    mo = O.build_lgssm(O.Matern32(), O.RegularSpacing(0.0, 0.1, T), 0.1)
    y = O.sample_prior(mo, np.random.default_rng(20261017 + 1))
    fx = pkg.to_sde(pkg.GP(pkg.Matern32Kernel()))(pkg.RegularSpacing(0.0, 0.1, T), 0.1)
    cm = c_oracle.Model.from_lgssm(mo)
    t_gpu, lml = timeit(lambda: pkg.gp.logpdf(fx, y), 20)
    t_cpu, ref = timeit(lambda: c_oracle.logpdf(cm, y), 20)
    return {"config": "cfg1 Matern32 D=2 T=10000 sigma2=0.1 logpdf (README example)", "gpu_ms_e2e": t_gpu * 1e3, "cpu_oracle_ms": t_cpu * 1e3,
            "lml_rel_err": abs(lml - ref) / abs(ref)}


def cfg3(T):
    TK = pkg.gp.TransformedKernel
    kp = 1.0 * pkg.Matern32Kernel() + 0.7 * pkg.Matern52Kernel() + 0.5 * TK(pkg.Matern52Kernel(), 0.5) + 0.3 * TK(pkg.Matern32Kernel(), 2.0)
    ko = 1.0 * O.Matern32() + 0.7 * O.Matern52() + 0.5 * O.Matern52().stretch(0.5) + 0.3 * O.Matern32().stretch(2.0)
    mo = O.build_lgssm(ko, O.RegularSpacing(0.0, 0.01, T), 0.1)
    rng = np.random.default_rng(20261017 + 3)
    y = np.sin(np.arange(T) * 0.004) + 0.3 * np.cos(np.arange(T) * 0.05) + 0.35 * rng.standard_normal(T)
    fx = pkg.to_sde(pkg.GP(kp))(pkg.RegularSpacing(0.0, 0.01, T), 0.1)
    model = fx.build_lgssm()
    t_gpu, (mu, var) = timeit(lambda: pkg.lgssm.posterior_marginals(model, y, 1e-2, h), 5, 2)
    # device-resident inputs / outputs: the library's own time without the PCIe copies
    mm = pkg.lgssm._Marshalled(model)
    yd = torch.from_numpy(y).cuda(); Rn = torch.full((1,), 1e-2, dtype=torch.float64, device="cuda")
    md = torch.empty(T, dtype=torch.float64, device="cuda"); vd = torch.empty(T, dtype=torch.float64, device="cuda")
    t_dev, _ = timeit(lambda: h.posterior_marginals(mm.desc, yd, Rn, 0, md, vd, None), 10, 3)
    h.set_algo(pkg.TGP_ALGO_SCAN)
    t_dev_general, _ = timeit(lambda: h.posterior_marginals(mm.desc, yd, Rn, 0, md, vd, None), 3, 1)
    h.set_algo(pkg.TGP_ALGO_AUTO)
    h.set_timing(True)
    for _ in range(3):
        h.posterior_marginals(mm.desc, yd, Rn, 0, md, vd, None)
    tim = h.timing()
    h.set_timing(False)
    t_lp, lml = timeit(lambda: pkg.lgssm.logpdf(model, y, h), 3, 1)
    cm = c_oracle.Model.from_lgssm(mo)
    t0 = time.perf_counter()
    mu_o, var_o, lml_o = c_oracle.posterior_marginals(cm, y, 1e-2)
    t_cpu = time.perf_counter() - t0
    return {"config": f"cfg3 D=10 sum(Matern32+Matern52+Matern52∘ST(.5)+Matern32∘ST(2)) T={T} posterior marginals (filter + RTS)",
            "gpu_ms_e2e_host_buffers": t_gpu * 1e3, "gpu_ms_device_resident": t_dev * 1e3, "gpu_steps_per_s": T / t_dev,
            "gpu_ms_device_resident_general_scan": t_dev_general * 1e3, "hbm_frac_of_measured": 1784.0 * T / t_dev / 1e9 / 6552.3,
            "gpu_logpdf_ms": t_lp * 1e3,
            "cpu_oracle_ms": t_cpu * 1e3, "cpu_steps_per_s": T / t_cpu,
            "mean_max_abs_err": float(np.max(np.abs(mu - mu_o))), "var_max_rel_err": float(np.max(np.abs(var - var_o) / var_o)),
            "lml_rel_err": abs(lml - lml_o) / abs(lml_o),
            "kernels_ms_per_call": {n: ms / 3 for n, ms, c in tim}}


def cfg5(Nr, T, T_cpu):
    """Separable(SE, Matern52) on RectilinearGrid(range(-3, 3, Nr), RegularSpacing(0, 0.01, T)): D = 3 Nr, M = Nr. FP64, dense
    as the reference executes it (2.57 GFLOP/step at Nr = 256). The CPU side is the NumPy oracle (multi-threaded BLAS) on T_cpu steps."""
    r = np.linspace(-3.0, 3.0, Nr)
    rng = np.random.default_rng(20261017 + 5)
    y = rng.standard_normal((T, Nr))
    fx = pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern52Kernel())))(pkg.RectilinearGrid(r, pkg.RegularSpacing(0.0, 0.01, T)), 0.1)
    model = fx.build_lgssm()
    t_gpu, lml = timeit(lambda: pkg.lgssm.logpdf(model, y, h), 2, 1)
    mo = O.build_lgssm_separable(O.SqExp(), O.Matern52(), r, O.RegularSpacing(0.0, 0.01, T_cpu), 0.1)
    t0 = time.perf_counter()
    ref = O.logpdf_steps(mo, y[:T_cpu])
    t_cpu = (time.perf_counter() - t0) / T_cpu
    lml_c, steps = pkg.lgssm.logpdf(pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern52Kernel())))(
        pkg.RectilinearGrid(r, pkg.RegularSpacing(0.0, 0.01, T_cpu)), 0.1).build_lgssm(), y[:T_cpu], h, per_step=True)
    D, M = 3 * Nr, Nr
    flops = 4.0 * D ** 3 + 2.0 * M * D * D + 2.0 * M * M * D + M ** 3 / 3.0 + 1.0 * M * M * D + 2.0 * D * D * M
    return {"config": f"cfg5 Separable(SE, Matern52) Nr={Nr} (D={D}, M={M}) T={T} logpdf, FP64, dense (library GEMMs + own Cholesky), CUDA-graph replay",
            "gpu_ms_per_step": t_gpu / T * 1e3, "gpu_steps_per_s": T / t_gpu, "gpu_tflops_dense_equiv": flops * T / t_gpu / 1e12,
            "cpu_oracle_ms_per_step": t_cpu * 1e3, "cpu_sample_steps": T_cpu, "lml_step_max_rel_err": float(np.max(np.abs(steps - ref) / np.abs(ref)))}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--T3", type=int, default=1_000_000)
    a = ap.parse_args()
    print(json.dumps(cfg1()))
    print(json.dumps(cfg3(a.T3)))
    # config 5: tools/cfg5_bench.py (full size, both arithmetic variants)
