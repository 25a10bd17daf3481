cd /root/repo
cat > /tmp/worker2.py <<'PY'
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
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
T = int(os.environ["TGP_T"])
b = sharded.shard_bounds(T, world); lo, hi = b[rank], b[rank+1]
h = pkg.Handle(0)
mm = pkg.lgssm._Marshalled(pkg.to_sde(pkg.GP(pkg.Matern52Kernel()))(pkg.RegularSpacing(0.0, 0.01, hi-lo), 0.1).build_lgssm())
sh = sharded.ShardedLogpdf(h, mm, rank, world, dev, dist, route="fir")
cm = c_oracle.Model.from_lgssm(O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, 0.01, T), 0.1))
out = torch.zeros(1, dtype=torch.float64, device=dev)
rng = np.random.default_rng(5)
y = np.sin(np.arange(T) * 0.003) + 0.4 * rng.standard_normal(T)
yd = torch.from_numpy(np.ascontiguousarray(y[lo:hi])).to(dev)
ref = c_oracle.logpdf(cm, y)
for rep in range(3):
    sh.logpdf(yd, out); torch.cuda.synchronize()
    print("rank", rank, "T", T, "rel", abs(float(out.item()) - ref) / abs(ref), flush=True)
for rep in range(6):
    sh.logpdf(yd, None, sync=False)
sh.result(out); sh.check(); torch.cuda.synchronize()
print("rank", rank, "pipelined rel", abs(float(out.item()) - ref) / abs(ref), flush=True)
dist.barrier(); dist.destroy_process_group()
PY
for T in ${TS:-4000000 20000000}; do TGP_ROOT=/root/repo TGP_T=$T OMP_NUM_THREADS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node=${NP:-2} --master-addr 127.0.0.1 --master-port 29541 /tmp/worker2.py 2>&1 | grep -v Warning | grep "rel\|Error\|error" | head -12; done
