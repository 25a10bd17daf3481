"""
build.py — compiles libtgpb200.so (the C-ABI library, include/tgp_b200.h) for sm_100a, in-tree.

One nvcc object per latent dimension (tgp_inst.cu with -DTGP_D=<D>) plus the ABI translation unit,
compiled in parallel and linked into temporalgps.jl_b200/libtgpb200.so. Objects are cached under
temporalgps.jl_b200/build/ keyed by a hash of the sources and flags, so an unchanged tree is a no-op.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libtgpb200.so")

# keep in step with TGP_FOR_EACH_D in csrc/tgp_dispatch.h
TGP_DIMS = (1, 2, 3, 4, 5, 6, 8, 10)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
] + os.environ.get("TGP_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libtgpb200.so cannot be built (there is no CPU fallback)")


def _source_hash() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(repr(TGP_DIMS).encode())
    return h.hexdigest()[:16]


def _compile(args):
    src, obj, extra, log = args
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {os.path.basename(obj)}:\n{p.stderr[-4000:]}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if needed) and return the path of libtgpb200.so."""
    os.makedirs(BUILD, exist_ok=True)
    tag = _source_hash()
    stamp = os.path.join(BUILD, "stamp")
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == tag:
        return LIB
    jobs = [(os.path.join(CSRC, "tgp_api.cu"), os.path.join(BUILD, "tgp_api.o"), [], os.path.join(BUILD, "tgp_api.log"))]
    jobs.append((os.path.join(CSRC, "tgp_dense.cu"), os.path.join(BUILD, "tgp_dense.o"), [], os.path.join(BUILD, "tgp_dense.log")))
    jobs.append((os.path.join(CSRC, "tgp_xchg.cu"), os.path.join(BUILD, "tgp_xchg.o"), [], os.path.join(BUILD, "tgp_xchg.log")))
    jobs.append((os.path.join(CSRC, "tgp_seq.cu"), os.path.join(BUILD, "tgp_seq.o"), [], os.path.join(BUILD, "tgp_seq.log")))
    jobs.append((os.path.join(CSRC, "tgp_lti.cu"), os.path.join(BUILD, "tgp_lti.o"), [], os.path.join(BUILD, "tgp_lti.log")))
    for d in TGP_DIMS:
        jobs.append((os.path.join(CSRC, "tgp_inst.cu"), os.path.join(BUILD, f"tgp_inst_d{d}.o"), [f"-DTGP_D={d}"],
                     os.path.join(BUILD, f"tgp_inst_d{d}.log")))
    with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(_compile, jobs))
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "lib64")
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-lcudart", "-L" + cuda_lib, "-Xlinker", "-rpath=" + cuda_lib]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("link failed:\n" + p.stderr[-4000:])
    with open(stamp, "w") as fh:
        fh.write(tag)
    if verbose:
        print(f"built {LIB}", file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
