import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
pkg = g.load_package()
h = pkg.default_handle(0)
out = torch.zeros(1, dtype=torch.float64, device="cuda")
for T in (65_536, 10_000_000):
    mm = pkg.lgssm._Marshalled(pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 0.01, T), 0.1).build_lgssm())
    ys = [torch.from_numpy(np.random.default_rng(i).standard_normal(T)).cuda() for i in range(3)]
    for i in range(4):
        h.logpdf(mm.desc, ys[i % 3], out)
