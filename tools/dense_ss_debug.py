import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
for Nr, T in ((16, 1800), (64, 1500), (256, 1200)):
    r = np.linspace(-3.0, 3.0, Nr)
    fx = pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern52Kernel())))(pkg.RectilinearGrid(r, pkg.RegularSpacing(0.0, 0.01, T)), 0.1)
    y = np.random.default_rng(1).standard_normal((T, Nr))
    print(Nr, pkg.lgssm.logpdf(fx.build_lgssm(), y, fx._handle()), flush=True)
