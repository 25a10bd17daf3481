"""World-size-2 test of the time-sharded logpdf's HOST logic on CPU (gloo): shard bounds, the all-gather of one scan
element per rank, the prefix fold (tgp_shard_prefix, the product's host function) and the final all-reduce. The
per-shard device kernels are replaced by CPU stand-ins (tests/emul for the shard element, the oracle for the shard's
log-likelihood) — this test checks the plumbing, the GPU tests check the kernels."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, ctypes as C
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["TGP_ROOT"])
import __graft_entry__ as g
from oracle import c_oracle, tgp_oracle as O
from tests.util import random_lgssm, sample_y
pkg = g.load_package()
from temporalgps_jl_b200 import sharded

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
emul = C.CDLL(os.environ["TGP_EMUL"])
rng = np.random.default_rng(42)          # same series on every rank
T, D = 1001, 3
m = random_lgssm(rng, T, D, "forward", True)
y = sample_y(rng, m)
b = sharded.shard_bounds(T, world)
s, e = b[rank], b[rank + 1]
sm = O.LGSSM("forward", m.As[s:e], m.as_[s:e], m.Qs[s:e], m.m0, m.P0, m.Hs[s:e], m.hs[s:e], m.Rs[s:e])
cm = c_oracle.Model.from_lgssm(sm)
ES = 3 * D * D + 2 * D
elem = np.zeros(ES)
ys = np.ascontiguousarray(y[s:e])
assert emul.emul_shard_reduce(C.byref(cm.desc), ys.ctypes.data_as(C.c_void_p), elem.ctypes.data_as(C.c_void_p)) == 0
gathered = torch.zeros(world * ES, dtype=torch.float64)
dist.all_gather_into_tensor(gathered, torch.from_numpy(elem))
elems = gathered.numpy().reshape(world, ES)
m_in, P_in = sharded.incoming_state(pkg._lib.shard_prefix, D, elems, rank, m.m0, m.P0)
sm2 = O.LGSSM("forward", sm.As, sm.as_, sm.Qs, m_in, P_in, sm.Hs, sm.hs, sm.Rs)
part = torch.tensor([O.logpdf(sm2, ys)], dtype=torch.float64)
dist.all_reduce(part)
ref = O.logpdf(m, y)
assert abs(part.item() - ref) <= 1e-9 * abs(ref), (part.item(), ref)
# the incoming state must equal the sequential filter's state at the shard boundary
if rank > 0:
    ms, Ps, _ = O.filter_(m, y)
    np.testing.assert_allclose(m_in, ms[s - 1], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(P_in, Ps[s - 1], rtol=1e-9, atol=1e-12)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_shard_bounds(pkg):
    from temporalgps_jl_b200 import sharded
    assert sharded.shard_bounds(10, 3) == [0, 4, 7, 10]
    assert sharded.shard_bounds(8, 8) == list(range(9))
    b = sharded.shard_bounds(80_000_000, 8)
    assert b[-1] == 80_000_000 and all(b[i + 1] - b[i] == 10_000_000 for i in range(8))


def test_time_sharded_logpdf_world2_gloo(pkg, tmp_path):
    so = os.path.join(ROOT, "tests", "emul", "_build", "libemul_scan.so")
    src = os.path.join(ROOT, "tests", "emul", "emul_scan.cpp")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", src, "-o", so], check=True)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TGP_ROOT=ROOT, TGP_EMUL=so, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert p.stdout.count("ok") == 2
