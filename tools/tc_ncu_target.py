#!/usr/bin/env python
"""Target for ncu: a few full dense steps of the config-5 shape on the tensor-core path, launched directly (timing mode
disables the CUDA-graph replay so every kernel is an ordinary launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
Nr, T = 256, 6
r = np.linspace(-3.0, 3.0, Nr)
fx = pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern52Kernel())), pkg.ArrayStorage(np.float32))(
    pkg.RectilinearGrid(r, pkg.RegularSpacing(0.0, 0.01, T)), 0.1)
y = np.random.default_rng(1).standard_normal((T, Nr))
h = fx._handle()
h.set_timing(True)
print(pkg.lgssm.logpdf(fx.build_lgssm(), y, h))
