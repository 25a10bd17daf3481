#!/usr/bin/env python
"""Target for ncu: config-5 shape, T = 20 000, FP32 storage: transient (dense steps) + time-blocked steady phase."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
Nr, T = 256, int(os.environ.get("TGP_T", "20000"))
r = np.linspace(-3.0, 3.0, Nr)
fx = pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern52Kernel())), pkg.ArrayStorage(np.float32))(
    pkg.RectilinearGrid(r, pkg.RegularSpacing(0.0, 0.01, T)), 0.1)
y = np.random.default_rng(1).standard_normal((T, Nr))
h = fx._handle()
t0 = time.perf_counter()
print(pkg.lgssm.logpdf(fx.build_lgssm(), y, h), time.perf_counter() - t0)
