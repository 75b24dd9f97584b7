#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS dump with `nvdisasm -g` line info and aggregate the
executed warp-instructions / stall samples per CUDA source line (innermost inlined frame).

usage: sass_lines.py <src.csv from ncu> <nvdisasm -g -c output> <mangled kernel name substring> [top N]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# 1. per-instruction metrics from ncu (first kernel instance in the file)
rows = list(csv.reader(open(src_csv)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
inst = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name" or r[0] == "Address":
        break
    inst.append(r)

# 2. line info per instruction from nvdisasm (in address order inside the kernel's section)
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
cur = ("?", 0)
per_inst = []
for l in lines[start + 1:]:
    if l.startswith("//---------------------") or l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        per_inst.append(cur)
print(f"ncu instructions: {len(inst)}  nvdisasm instructions: {len(per_inst)}")
n = min(len(inst), len(per_inst))
agg = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r, key in zip(inst[:n], per_inst[:n]):
    ex = int(r[col["Instructions Executed"]] or 0)
    th = int(r[col["Thread Instructions Executed"]] or 0)
    sm = int(r[col["# Samples"]] or 0)
    for a, v in zip((agg[key], tot), ((ex, th, sm),) * 2):
        a[0] += v[0]; a[1] += v[1]; a[2] += v[2]
print(f"total warp-inst {tot[0]:,}  thread-inst {tot[1]:,}  samples {tot[2]:,}")
print(f"{'file:line':32s} {'warp-inst':>12s} {'%':>6s} {'thr/inst':>8s} {'samples%':>8s}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0] + ':' + str(key[1]):32s} {v[0]:12,d} {100 * v[0] / tot[0]:6.2f} {v[1] / max(v[0], 1):8.1f} "
          f"{100 * v[2] / max(tot[2], 1):8.2f}")
