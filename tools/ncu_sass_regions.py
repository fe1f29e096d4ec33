#!/usr/bin/env python
"""Bucket the stall samples of an ncu SASS source dump by code region (150-instruction buckets)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; body = rows[2:]
iS = hdr.index("# Samples"); iSrc = hdr.index("Source"); iEx = hdr.index("Instructions Executed")
stall = [i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in body); print("total", tot, "n", len(body))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 150
for b0 in range(0, len(body), B):
    blk = body[b0:b0+B]; s = sum(int(r[iS]) for r in blk)
    if s < tot*0.005: continue
    ops = collections.Counter(); st = collections.Counter(); ex = 0
    for r in blk:
        op = r[iSrc].split()[0] if not r[iSrc].strip().startswith("@") else r[iSrc].split()[1]
        ops[op.split('.')[0]] += int(r[iS]); ex = max(ex, int(r[iEx]))
        for i in stall: st[hdr[i]] += int(r[i])
    print(b0, s, round(100*s/tot,1), "maxexec", ex, ops.most_common(4), st.most_common(3))
