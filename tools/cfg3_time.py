import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
h = pkg.default_handle(0)
T3 = 1_000_000
TK = pkg.gp.TransformedKernel
k3 = 1.0 * pkg.Matern32Kernel() + 0.7 * pkg.Matern52Kernel() + 0.5 * TK(pkg.Matern52Kernel(), 0.5) + 0.3 * TK(pkg.Matern32Kernel(), 2.0)
m3 = pkg.lgssm._Marshalled(pkg.to_sde(pkg.GP(k3))(pkg.RegularSpacing(0.0, 0.01, T3), 0.1).build_lgssm())
rng = np.random.default_rng(20261017 + 3)
y3 = torch.from_numpy(np.sin(np.arange(T3) * 0.004) + 0.3 * np.cos(np.arange(T3) * 0.05) + 0.35 * rng.standard_normal(T3)).cuda()
Rn = torch.full((1,), 1e-2, dtype=torch.float64, device="cuda")
md = torch.empty(T3, dtype=torch.float64, device="cuda"); vd = torch.empty(T3, dtype=torch.float64, device="cuda")
for _ in range(3): h.posterior_marginals(m3.desc, y3, Rn, 0, md, vd, None)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): h.posterior_marginals(m3.desc, y3, Rn, 0, md, vd, None)
torch.cuda.synchronize(); print("cfg3 ms", (time.perf_counter() - t0) / 10 * 1e3)
h.set_timing(True); h.posterior_marginals(m3.desc, y3, Rn, 0, md, vd, None); print([(n, round(ms / c * 1e3, 1)) for n, ms, c in h.timing()][:8])
