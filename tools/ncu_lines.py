#!/usr/bin/env python
"""Attribute the per-instruction counters of an ncu report to CUDA source lines.
usage: ncu_lines.py <ncu --page source --csv file> <object file .o> <kernel mangled-name substring> [top N]
Joins the SASS listing of the report with `nvdisasm --print-line-info` of the same cubin by instruction order."""
import collections, csv, os, re, subprocess, sys, tempfile
src_csv, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv, errors="replace")))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[0]
hdr, data = b["rows"][0], b["rows"][1:]
col = {n: i for i, n in enumerate(hdr)}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
# isolate the kernel's section
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
sass, cur_line = [], None
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        if sass: break
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        sass.append((cur_line, m.group(2)))
print(f"kernel {b['name'][:80]}: report {len(data)} instr, cubin {len(sass)} instr")
n = min(len(data), len(sass))
agg = collections.defaultdict(lambda: [0.0, 0.0])
T = S = 0.0
for i in range(n):
    ex = float(data[i][col["Instructions Executed"]] or 0); sm = float(data[i][col["# Samples"]] or 0)
    agg[sass[i][0]][0] += ex; agg[sass[i][0]][1] += sm; T += ex; S += sm
srcs = {}
def text(k):
    if k is None: return "?"
    f, ln = k
    if f not in srcs:
        for root in ("rs_image_segmentation_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p): srcs[f] = open(p).read().splitlines(); break
        else: srcs[f] = []
    return srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""
print("  inst%  samp%  file:line  source")
for k, (ex, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {100*ex/T:5.1f}  {100*sm/S:5.1f}  {(k[0] + ':' + str(k[1])) if k else '?':<24} {text(k)}")
