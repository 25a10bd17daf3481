#!/usr/bin/env python
"""Target for ncu: config 3 (D = 10, T = 1e6) posterior marginals through the steady smoother, device-resident buffers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g
pkg = g.load_package()
h = pkg.default_handle(0)
T = 1_000_000
TK = pkg.gp.TransformedKernel
k3 = 1.0 * pkg.Matern32Kernel() + 0.7 * pkg.Matern52Kernel() + 0.5 * TK(pkg.Matern52Kernel(), 0.5) + 0.3 * TK(pkg.Matern32Kernel(), 2.0)
mm = pkg.lgssm._Marshalled(pkg.to_sde(pkg.GP(k3))(pkg.RegularSpacing(0.0, 0.01, T), 0.1).build_lgssm())
rng = np.random.default_rng(3)
y = torch.from_numpy(np.sin(np.arange(T) * 0.004) + 0.35 * rng.standard_normal(T)).cuda()
Rn = torch.full((1,), 1e-2, dtype=torch.float64, device="cuda")
md = torch.empty(T, dtype=torch.float64, device="cuda")
vd = torch.empty(T, dtype=torch.float64, device="cuda")
for _ in range(3):
    h.posterior_marginals(mm.desc, y, Rn, 0, md, vd, None)
torch.cuda.synchronize()
print(float(md[0]), float(vd[0]))
