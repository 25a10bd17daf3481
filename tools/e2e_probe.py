"""Diagnostic: where does the end-to-end time of a host-buffer logpdf call go? (GPU box only)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
import bench

pkg = g.load_package()
T = 10_000_000
y = bench.synth_y(T, 1)
pin = torch.from_numpy(y).pin_memory()
pin_np = pin.numpy()
dev = torch.empty(T, dtype=torch.float64, device="cuda")
h = pkg.default_handle(0)
fx = pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 0.01, T), 0.1)
mm = pkg.lgssm._Marshalled(fx.build_lgssm())
out = np.zeros(1)

def timeit(name, fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    print(f"{name:40s} {(time.perf_counter() - t0) / n * 1e3:8.3f} ms")

timeit("torch H2D pinned 80MB", lambda: dev.copy_(pin, non_blocking=True))
timeit("torch H2D pageable 80MB", lambda: dev.copy_(torch.from_numpy(y)))
timeit("h.logpdf(device y)", lambda: h.logpdf(mm.desc, dev, out))
timeit("h.logpdf(pinned host y)", lambda: h.logpdf(mm.desc, pin_np, out))
timeit("h.logpdf(pageable host y)", lambda: h.logpdf(mm.desc, y, out))
timeit("gp.logpdf(pinned host y)", lambda: pkg.gp.logpdf(fx, pin_np))
timeit("build_lgssm + marshal", lambda: pkg.lgssm._Marshalled(fx.build_lgssm()))
h.set_algo(pkg.TGP_ALGO_SCAN)
timeit("h.logpdf(device y) algo=scan", lambda: h.logpdf(mm.desc, dev, out))
print(h.counters())
