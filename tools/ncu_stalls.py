#!/usr/bin/env python
"""Aggregate the per-instruction page of an ncu report (`ncu -i X --page source --csv`) : stall reasons in total,
instruction mix by opcode, and the hottest instructions."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
# several kernels may be concatenated: take the first block unless an index is given
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
b = blocks[which]
hdr = b["rows"][0]; data = b["rows"][1:]
col = {n: i for i, n in enumerate(hdr)}
print("kernel:", b["name"][:120], " instructions:", len(data))
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = collections.Counter()
for r in data:
    for s in stalls:
        tot[s] += float(r[col[s]] or 0)
S = sum(tot.values())
print("stall samples (all):", ", ".join(f"{k[6:]} {100*v/S:.1f}%" for k, v in tot.most_common(9)))
ops = collections.Counter(); opsamp = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += float(r[col["Instructions Executed"]] or 0)
    opsamp[op] += float(r[col["# Samples"]] or 0)
T = sum(ops.values())
print("warp instructions executed:", int(T))
print("mix:", ", ".join(f"{k} {100*v/T:.1f}%" for k, v in ops.most_common(18)))
TS = sum(opsamp.values())
print("samples by op:", ", ".join(f"{k} {100*v/TS:.1f}%" for k, v in opsamp.most_common(12)))
print("hottest instructions:")
for r in sorted(data, key=lambda r: -float(r[col["# Samples"]] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
    top = max(stalls, key=lambda s: float(r[col[s]] or 0))
    print(f"  {r[col['# Samples']]:>7} {top[6:]:<12} {r[col['Source']].strip()[:90]}")
