#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` dump: stall samples by opcode and the hottest lines."""
import csv, sys, collections
path = sys.argv[1]; kernel_idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
# split per kernel
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = r; continue
    cur["rows"].append(r)
b = blocks[kernel_idx]; hdr = b["hdr"]
iS = hdr.index("# Samples"); iSrc = hdr.index("Source"); iEx = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in b["rows"])
print(b["name"], "total samples", tot, "instructions", len(b["rows"]))
byop = collections.Counter(); exop = collections.Counter()
for r in b["rows"]:
    op = r[iSrc].split()[0] if not r[iSrc].strip().startswith("@") else r[iSrc].split()[1]
    byop[op.split(".")[0]] += int(r[iS]); exop[op.split(".")[0]] += int(r[iEx])
print("samples by opcode:", [(k, v, round(100*v/tot,1)) for k, v in byop.most_common(14)])
print("executed by opcode:", exop.most_common(14))
st = collections.Counter()
for r in b["rows"]:
    for i in stall_cols: st[hdr[i]] += int(r[i])
print("stall reasons:", [(k, v, round(100*v/tot,1)) for k, v in st.most_common(8)])
top = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][iS]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]
for i in sorted(top):
    r = b["rows"][i]
    reasons = sorted(((int(r[j]), hdr[j]) for j in stall_cols if int(r[j])), reverse=True)[:3]
    print(i, r[iS], r[iSrc].strip()[:70], reasons)
