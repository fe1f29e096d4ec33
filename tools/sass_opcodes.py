#!/usr/bin/env python
"""SASS opcode histogram per kernel of the shipped library:  python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "dlsa_b200", "libdlsa_b200.so")], capture_output=True, text=True).stdout
kern, hist = None, collections.defaultdict(collections.Counter)
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
keys = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DMUL", "DADD", "LDS", "STS", "REDG", "BAR", "HMMA", "UTCHMMA", "LDTM", "UTMALDG"]
names = list(hist)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass dlsa_b200/libdlsa_b200.so: static instruction counts per sm_100a kernel")
print("# DMMA = mma.sync.m8n8k4.f64 (the FP64 tensor path: tcgen05 has no f64 kind), UBLKCP = cp.async.bulk (TMA, 1-D),")
print("# SYNCS = mbarrier, LDGSTS = cp.async, REDG = red.global (fixed-point pair histogram); no HMMA / UTC*MMA / LDTM by design")
print("kernel,total," + ",".join(keys))
tot = collections.Counter()
for k, d in sorted(zip(names, dem), key=lambda kv: -sum(hist[kv[0]].values())):
    h = hist[k]
    print('"' + d[:120] + '",' + str(sum(h.values())) + "," + ",".join(str(h.get(x, 0)) for x in keys))
    for x in keys:
        tot[x] += h.get(x, 0)
print("TOTAL," + str(sum(sum(h.values()) for h in hist.values())) + "," + ",".join(str(tot[x]) for x in keys))
