#!/usr/bin/env python
"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view): prints the hottest
instructions with their dominant stall reasons, plus an opcode histogram of executed instructions."""
import csv
import collections
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    data.append(r)
tot = sum(int(r[idx["# Samples"]]) for r in data)
print("total samples", tot)
ops = collections.Counter()
for r in data:
    op = r[idx["Source"]].split()
    if not op:
        continue
    name = op[1] if op[0].startswith("@") else op[0]
    ops[name.split(".")[0]] += int(r[idx["Instructions Executed"]])
print("executed warp-instructions by opcode:", ", ".join(f"{k}:{v/1e6:.2f}M" for k, v in ops.most_common(18)))
top = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[idx[c]]), c) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {int(r[idx['# Samples']]):6d} {100.0*int(r[idx['# Samples']])/tot:5.1f}%  {r[idx['Source']][:70]:70s} " +
          " ".join(f"{c[6:]}={n}" for n, c in st if n))
