#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel in an ncu report (needs -lineinfo +
--import-source on).  Usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import io
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# find header row
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
si, ii, wi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
ai = hdr.index("Address") if "Address" in hdr else None
li = hdr.index("#") if "#" in hdr else None
agg = {}
cur = None
tot_i = tot_s = 0
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    src = r[si]
    is_sass = ai is not None and r[ai].startswith("0x")
    if not is_sass:
        cur = (r[li] if li is not None else "", src.strip())
        continue
    try:
        n, s = int(r[ii] or 0), int(r[wi] or 0)
    except ValueError:
        continue
    a = agg.setdefault(cur, [0, 0])
    a[0] += n
    a[1] += s
    tot_i += n
    tot_s += s
print(f"total warp-instructions {tot_i}, stall samples {tot_s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[0]:>10} inst {100*v[0]/max(tot_i,1):5.1f}%  {v[1]:>7} smp {100*v[1]/max(tot_s,1):5.1f}%  | {k}")
