#!/usr/bin/env python
"""Per-FUNCTION attribution of a step-kernel ncu capture (companion of sass_lines.py).

usage: sass_funcs.py <src.csv from `ncu -i rep --page source --csv`> <`nvdisasm -g -c` of vss_step cubin>
                     <mangled kernel name substring> [tiles per launch, default 32768] [top N]

Joins the per-instruction counters with nvdisasm's line info (innermost inlined frame) and sums
executed warp-instructions, active lanes and stall samples per function of csrc/vss_lane.cuh.
"""
import csv
import os
import re
import sys
from collections import defaultdict

src_csv, dis, kname = sys.argv[1:4]
tiles = int(sys.argv[4]) if len(sys.argv) > 4 else 32768
top = int(sys.argv[5]) if len(sys.argv) > 5 else 45
here = os.path.dirname(os.path.abspath(__file__))
lane_src = open(os.path.join(here, "..", "rsoccer_isaac_cleanrl_b200", "csrc", "vss_lane.cuh")).read().split("\n")

rows = list(csv.reader(open(src_csv)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {h: i for i, h in enumerate(rows[hdr_i])}
inst = [r for r in rows[hdr_i + 1:] if r and r[0] not in ("Kernel Name", "Address")]

lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
cur, per = ("?", 0), []
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        per.append(cur)
assert len(per) == len(inst), (len(per), len(inst))

fn_at, fn = {}, "?"
for i, l in enumerate(lane_src, 1):
    m = re.match(r"^\s*(?:VSS_HD(?:_COLD)?|inline)\s+[\w:<>\*& ]+?\s+(\w+)\(", l)
    if m:
        fn = m.group(1)
    fn_at[i] = fn

agg, tot = defaultdict(lambda: [0, 0, 0, 0]), [0, 0, 0]
for r, key in zip(inst, per):
    ex, th, sm = (int(r[col[c]] or 0) for c in ("Instructions Executed", "Thread Instructions Executed", "# Samples"))
    k = fn_at.get(key[1], "?") if key[0] == "vss_lane.cuh" else (f"{key[0]}:{key[1]}" if key[0] == "vss_step.cu" else key[0])
    a = agg[k]
    a[0] += ex; a[1] += th; a[2] += sm; a[3] += 1
    tot[0] += ex; tot[1] += th; tot[2] += sm
print(f"total warp-inst {tot[0]:,} = {tot[0] / tiles:.0f} per 32-field tile; stall samples {tot[2]:,}; static {len(inst)}")
print(f"{'function':34s} {'inst/tile':>10s} {'%':>7s} {'lanes':>6s} {'samples%':>9s} {'static':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k:34s} {v[0] / tiles:10.1f} {100 * v[0] / tot[0]:7.2f} {v[1] / max(v[0], 1):6.1f} "
          f"{100 * v[2] / max(tot[2], 1):9.2f} {v[3]:7d}")
