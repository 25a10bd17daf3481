import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); h = pkg.default_handle(0)
Nr, T = 256, 12
r = np.linspace(-3.0, 3.0, Nr)
y = np.random.default_rng(0).standard_normal((T, Nr))
fx = pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern52Kernel())))(pkg.RectilinearGrid(r, pkg.RegularSpacing(0.0, 0.01, T)), 0.1)
m = fx.build_lgssm()
print(pkg.lgssm.logpdf(m, y, h))
print(pkg.lgssm.logpdf(m, y, h))
