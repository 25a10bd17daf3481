#!/usr/bin/env python
"""TF32 tensor-core rate of the library's OWN contraction kernel (k_tc_gemm_tn, tcgen05.mma kind::tf32 through TMA + TMEM) on large,
machine-filling products: the "measured (own)" denominator for the config-5 fractions (MEASURED_PEAKS.json has no TF32 figure).
A 3xTF32 product issues three MMAs per tile pair, so the MMA rate is 3 * 2 K M N / t; the useful (FP32-accurate) rate is 2 K M N / t.
    python tools/tc_peak.py            -> one JSON line"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

pkg = g.load_package()
h = pkg.Handle(0)
out = {"kernel": "k_tc_gemm_tn (3xTF32, FP32 accumulation in TMEM)", "shapes": []}
for K, M, N in [(2048, 2048, 2048), (4096, 4096, 4096), (1024, 8192, 4096)]:
    rng = np.random.default_rng(K)
    X = rng.standard_normal((K, M)).astype(np.float32)
    Y = rng.standard_normal((K, N)).astype(np.float32)
    h.tc_gemm(X[:, :256], Y[:, :256])                 # warm-up (descriptors, attributes)
    best = None
    for rep in range(3):
        h.set_timing(True)
        C = h.tc_gemm(X, Y)
        tim = [t for t in h.timing() if "gemm" in t[0] or "tc_gemm" in t[0]]
        h.set_timing(False)
        ms = max(t[1] / t[2] for t in tim) if tim else float("nan")
        best = ms if best is None else min(best, ms)
    err = float(np.max(np.abs(C[:64, :64] - X[:, :64].astype(np.float64).T @ Y[:, :64].astype(np.float64))))
    out["shapes"].append({"K": K, "M": M, "N": N, "ms": best, "mma_tflops": 3 * 2.0 * K * M * N / (best * 1e-3) / 1e12,
                          "useful_tflops": 2.0 * K * M * N / (best * 1e-3) / 1e12, "max_abs_err_64x64": err,
                          "timed": [t[0] for t in tim]})
out["tf32_mma_peak_measured_own_tflops"] = max(s["mma_tflops"] for s in out["shapes"])
print(json.dumps(out))
