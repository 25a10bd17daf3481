"""Kernel time of the one-launch logpdf as a function of T (events around each launch, no overlap between calls) and the host cost
of a call (Python + library, GPU idle): `python tools/fir_curve.py`."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
h = pkg.default_handle(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
h.set_stream(st.cuda_stream)
out = torch.zeros(1, dtype=torch.float64, device="cuda")
rng = np.random.default_rng(0)
big = [torch.from_numpy(rng.standard_normal(20_000_000)).cuda() for _ in range(4)]
for T in (65_536, 1_000_000, 2_000_000, 5_000_000, 10_000_000, 20_000_000):
    mm = pkg.lgssm._Marshalled(pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 0.01, T), 0.1).build_lgssm())
    ys = [b[:T] for b in big]
    for i in range(4):
        h.logpdf(mm.desc, ys[i % 4], out)
    h.synchronize()
    h.set_timing(True)
    for i in range(20):
        h.logpdf(mm.desc, ys[i % 4], out)
    tim = h.timing()
    h.set_timing(False)
    k_us = [ms / c * 1e3 for n, ms, c in tim if n.startswith("k_fir")][0]
    # back-to-back (PDL) rate
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200 if T <= 2_000_000 else 60
    h.synchronize()
    t0 = time.perf_counter()
    e0.record(st)
    for i in range(n):
        h.logpdf(mm.desc, ys[i % 4], out)
    e1.record(st)
    host_us = (time.perf_counter() - t0) / n * 1e6
    h.synchronize()
    print(f"T={T:9d} kernel alone {k_us:7.2f} us | back-to-back {e0.elapsed_time(e1) / n * 1e3:7.2f} us/call | host issue {host_us:6.2f} us/call", flush=True)
