#!/usr/bin/env python
"""Dynamic SASS profile of one kernel in an ncu report: executed warp-instructions by opcode, and the instructions
holding the most stall samples with their dominant stall reason.
Usage: python tools/ncu_sass.py report.ncu-rep kernel_regex [top_n] [--dump]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
dump = "--dump" in sys.argv
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
si, ii, wi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops, op_smp, reasons = Counter(), Counter(), Counter()
items = []
for n, r in enumerate(rows[hi + 1:]):
    if len(r) != len(hdr):
        continue
    try:
        ex, smp = int(r[ii] or 0), int(r[wi] or 0)
    except ValueError:
        continue
    src = r[si].strip()
    t = src.split()
    op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
    op = op.split(".")[0].rstrip(";")
    ops[op] += ex
    op_smp[op] += smp
    st = {h: int(r[i] or 0) for i, h in stall_cols}
    for h, v in st.items():
        reasons[h] += v
    dom = max(st.items(), key=lambda kv: kv[1]) if st else ("", 0)
    items.append((n, src, ex, smp, dom))
tot_i, tot_s = sum(ops.values()), sum(op_smp.values())
print(f"kernel {kern}: {tot_i} warp-instructions executed, {tot_s} stall samples")
print("stall reasons:", ", ".join(f"{h[6:]} {100*v/max(tot_s,1):.1f}%" for h, v in reasons.most_common(8)))
print("by opcode (executed, % / samples %):")
for op, c in ops.most_common(24):
    print(f"  {op:10s} {c:>10} {100*c/tot_i:5.1f}%   {100*op_smp[op]/max(tot_s,1):5.1f}%")
print(f"top {top} instructions by stall samples:")
for n, src, ex, smp, dom in sorted(items, key=lambda x: -x[3])[:top]:
    print(f"  #{n:5d} {smp:6d} smp {100*smp/max(tot_s,1):4.1f}%  ex {ex:>8}  {dom[0][6:]:12s} {src[:90]}")
if dump:
    for n, src, ex, smp, dom in items:
        print(f"#{n:5d} ex {ex:>8} smp {smp:5d} {dom[0][6:] if smp else '':12s} {src}")
