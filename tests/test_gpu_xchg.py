"""GPU test of the time-sharded logpdf: TWO processes, both on cuda:0 (the test box has one GPU; CUDA IPC works between
processes on the same device), rendezvous over gloo. Each rank owns a shard of one series (uneven shards included).
route "fir": tgp_shard_logpdf — one launch per shard, the halo and the partial log-likelihoods travel through the mapped peer
buffers + flags (tgp_xchg.cu), no collective. route "steady": tgp_shard_phase1 / phase2 around the caller's all-gather.
The total must equal the sequential oracle on the whole series (1e-6 relative, north_star) for several consecutive calls
(epoch ring, slot reuse), synchronised and pipelined."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["TGP_ROOT"])
import __graft_entry__ as g
from oracle import c_oracle, tgp_oracle as O
pkg = g.load_package()
from temporalgps_jl_b200 import sharded

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
T = int(os.environ["TGP_T"])
b = sharded.shard_bounds(T, world)
lo, hi = b[rank], b[rank + 1]
Ts = hi - lo                                  # this rank's shard
h = pkg.Handle(0)
fx = pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 0.01, Ts), 0.1)
mm = pkg.lgssm._Marshalled(fx.build_lgssm())
ov = os.environ.get("TGP_OVERLAP") == "1"    # overlapped scatter: every shard carries the 3072 observations before it
sh = sharded.ShardedLogpdf(h, mm, rank, world, dev, dist, route=os.environ["TGP_ROUTE"], overlap=ov)
keep = []
def put(y):
    if ov:
        buf, view = sharded.shard_with_halo(torch, y, b, rank, dev)
        keep.append(buf)
        return view
    return torch.from_numpy(np.ascontiguousarray(y[lo:hi])).to(dev)
assert sh.route == os.environ["TGP_ROUTE"], (sh.route, getattr(sh, "transport_error", None))
mo = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.01, T), 0.1)
cm = c_oracle.Model.from_lgssm(mo)
out = torch.zeros(1, dtype=torch.float64, device=dev)
for rep in range(4):
    rng = np.random.default_rng(100 + rep)    # same series on every rank
    y = np.sin(np.arange(T) * 0.003) + 0.4 * rng.standard_normal(T)
    yd = put(y)
    sh.logpdf(yd, out)
    torch.cuda.synchronize()
    ref = c_oracle.logpdf(cm, y)
    got = float(out.item())
    assert abs(got - ref) <= 1e-6 * abs(ref), (rep, got, ref)
# pipelined form: nothing waits on the host between calls; statuses accumulate on the device until check()
outs = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(3)]
refs = []
yds = []
for rep in range(3):
    rng = np.random.default_rng(200 + rep)
    y = np.cos(np.arange(T) * 0.002) + 0.4 * rng.standard_normal(T)
    refs.append(c_oracle.logpdf(cm, y))
    yds.append(put(y))
for rep in range(3):
    sh.logpdf(yds[rep], outs[rep], sync=False)
sh.check()
for rep in range(3):
    got = float(outs[rep].item())
    assert abs(got - refs[rep]) <= 1e-6 * abs(refs[rep]), (rep, got, refs[rep])
# both layouts interleaved on ONE handle (the exchange ring's acknowledgements must stay current while the overlap layout runs)
if os.environ["TGP_ROUTE"] == "fir":
    sh2 = sharded.ShardedLogpdf(h, mm, rank, world, dev, dist, route="fir", overlap=not ov)
    order = [sh, sh, sh, sh, sh, sh, sh2, sh2, sh, sh2, sh2, sh2, sh2, sh2, sh2, sh]
    outs = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in order]
    rng = np.random.default_rng(300)
    y = np.cos(np.arange(T) * 0.001) + 0.4 * rng.standard_normal(T)
    ref = c_oracle.logpdf(cm, y)
    buf, view = sharded.shard_with_halo(torch, y, b, rank, dev)
    for s_, o_ in zip(order, outs):
        s_.logpdf(view, o_, sync=False)
    sh.check()
    for i, o_ in enumerate(outs):
        got = float(o_.item())
        assert abs(got - ref) <= 1e-6 * abs(ref), ("mixed", i, got, ref)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("route,T", [("fir", 200_000), ("fir", 200_001), ("fir", 131_073), ("fir", 203_000), ("fir", 207_300), ("steady", 200_000),
                                     ("fir+overlap", 200_000), ("fir+overlap", 200_001)])
def test_time_sharded_logpdf_two_processes(pkg, tmp_path, route, T):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    route, _, ov = route.partition("+")
    env = dict(os.environ, TGP_ROOT=ROOT, OMP_NUM_THREADS="1", TGP_ROUTE=route, TGP_T=str(T), TGP_OVERLAP="1" if ov else "0")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert p.stdout.count("ok") == 2


def test_logpdf_device_outputs_are_only_enqueued(pkg):
    """tgp_logpdf with device-resident y and a device destination only ENQUEUES (one launch, nothing to report back: the plan
    checked positive-definiteness on the host); consecutive calls queue back to back — and may overlap (programmatic dependent
    launch) — and the results equal the synchronous calls."""
    import numpy as np
    import torch
    from oracle import c_oracle, tgp_oracle as O
    h = pkg.Handle(0)
    T = 200_000
    fx = pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 0.01, T), 0.1)
    mm = pkg.lgssm._Marshalled(fx.build_lgssm())
    cm = c_oracle.Model.from_lgssm(O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.01, T), 0.1))
    ys = [np.sin(np.arange(T) * 0.003 * (k + 1)) + 0.4 * np.random.default_rng(k).standard_normal(T) for k in range(6)]
    yd = [torch.from_numpy(y).cuda() for y in ys]
    outs = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in ys]
    c0 = h.counters()
    for y, o in zip(yd, outs):
        h.logpdf(mm.desc, y, o)
    c1 = h.counters()
    assert c1["launches"] - c0["launches"] == len(ys) and c1["d2h_bytes"] == c0["d2h_bytes"]
    h.synchronize()
    for y, o in zip(ys, outs):
        ref = c_oracle.logpdf(cm, y)
        assert abs(float(o.item()) - ref) <= 1e-11 * abs(ref)
    # a grid so fine that the plan declines: the call falls back to the two-phase / general kernels and still answers
    fx2 = pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 1e-6, T), 0.1)
    mm2 = pkg.lgssm._Marshalled(fx2.build_lgssm())
    cm2 = c_oracle.Model.from_lgssm(O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 1e-6, T), 0.1))
    out = np.zeros(1)
    h.logpdf(mm2.desc, yd[0], out)
    ref = c_oracle.logpdf(cm2, ys[0])
    assert abs(out[0] - ref) <= 1e-6 * abs(ref)
    h.logpdf(mm.desc, yd[1], outs[1])
    h.synchronize()
    ref = c_oracle.logpdf(cm, ys[1])
    assert abs(float(outs[1].item()) - ref) <= 1e-11 * abs(ref)
    h.close()
