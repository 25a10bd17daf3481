"""Driver for ncu captures of the config-2 logpdf call (device-resident y, T = 1e7): `ncu ... python tools/prof_fir.py [n_calls] [T]`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
h = pkg.default_handle(0)
mm = pkg.lgssm._Marshalled(pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 0.01, T), 0.1).build_lgssm())
rng = np.random.default_rng(0)
ys = [torch.from_numpy(rng.standard_normal(T)).cuda() for _ in range(5 if T <= 20_000_000 else 2)]
out = torch.zeros(1, dtype=torch.float64, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
h.set_stream(st.cuda_stream)
for i in range(3):
    h.logpdf(mm.desc, ys[i % len(ys)], out)
h.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(n):
    h.logpdf(mm.desc, ys[i % len(ys)], out)
e1.record(st)
h.synchronize()
torch.cuda.synchronize()
print("lml", float(out.item()), "launches", h.counters()["launches"], "us/call", e0.elapsed_time(e1) / n * 1e3,
      "stagger", os.environ.get("TGP_FIR_STAGGER"))
