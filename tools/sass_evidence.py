#!/usr/bin/env python
"""profiles/sass_rNN.txt: per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md):
UTCHMMA / UTCQMMA (tcgen05.mma), UTMALDG (TMA), LDTM / STTM (tcgen05.ld/st), LDGSTS (cp.async), DFMA (FP64), HMMA (legacy: must be 0).
Usage: python tools/sass_evidence.py > profiles/sass_r02.txt   (runs cuobjdump on the in-tree libtgpb200.so; no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "temporalgps.jl_b200", "libtgpb200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "LDGSTS", "DFMA", "DADD", "DMUL", "HMMA", "HGMMA",
        "LDG.E.128", "LDS.128", "SHFL", "BAR.SYNC", "ACQBULK", "MEMBAR", "NANOSLEEP"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    fn, cnt = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn)
            cnt[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cnt[fn]["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                cnt[fn][k] += 1
    print(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)} (sm_100a), instruction counts per kernel; legacy tensor paths (HMMA / HGMMA) must be 0")
    tot = collections.Counter()
    for fn, c in cnt.items():
        if c["total"] == 0:
            continue
        items = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
        print(f"{fn[:110]:110s} total={c['total']:6d} {items}")
        tot.update(c)
    print("# whole library: " + " ".join(f"{k}={tot[k]}" for k in KEYS))


if __name__ == "__main__":
    sys.exit(main())
