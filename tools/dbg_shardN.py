"""Multi-GPU diagnostic (torchrun, one rank per GPU): per-shard log-likelihoods of the one-launch sharded logpdf against the
oracle's conditionals, synchronised and pipelined."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
from oracle import c_oracle, tgp_oracle as O
pkg = g.load_package()
from temporalgps_jl_b200 import sharded
local = int(os.environ.get("LOCAL_RANK", "0"))
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
Ts = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
NB = 4
h = pkg.default_handle(local)
mm = pkg.lgssm._Marshalled(pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 0.01, Ts), 0.1).build_lgssm())
sh = sharded.ShardedLogpdf(h, mm, rank, world, dev)
assert sh.route == "fir", (sh.route, sh.transport_error)
def series(r, i):
    rng = np.random.default_rng(1000 * r + i)
    return np.sin(np.arange(Ts) * 0.003 + r) + 0.4 * rng.standard_normal(Ts)
ys = [torch.from_numpy(series(rank, i)).to(dev) for i in range(NB)]
refs = {}
if rank == 0:
    for i in range(NB):
        y = np.concatenate([series(r, i) for r in range(world)])
        cum = [0.0]
        for r in range(world):
            n = (r + 1) * Ts
            cum.append(c_oracle.logpdf(c_oracle.Model.from_lgssm(O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.01, n), 0.1)), y[:n]))
        refs[i] = np.diff(cum)
part = np.zeros(1); out = torch.zeros(1, dtype=torch.float64, device=dev)
def report(tag, i):
    h.shard_partial(part)
    allp = [None] * world
    dist.all_gather_object(allp, float(part[0]))
    if rank == 0:
        err = np.abs(np.array(allp) - refs[i]) / np.abs(refs[i])
        print(tag, "buf", i, "max rel err per shard", err.max(), "argmax", int(err.argmax()), "total rel", abs(float(out.item()) - refs[i].sum()) / abs(refs[i].sum()), flush=True)
for rep in range(6):
    sh.logpdf(ys[rep % NB], out); torch.cuda.synchronize()
    report("sync", rep % NB)
for rnd in range(3):
    for rep in range(25):
        sh.logpdf(ys[rep % NB], None, sync=False)
    sh.result(out); sh.check(); torch.cuda.synchronize()
    report("pipelined", 24 % NB)
    dist.barrier()
sh.logpdf(ys[0], out); torch.cuda.synchronize()
report("after", 0)
dist.barrier(); dist.destroy_process_group()
