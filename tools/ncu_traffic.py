#!/usr/bin/env python
"""Per-launch DRAM traffic of one kernel from an `ncu --set full` report -> a small JSON bench.py reads for
roofline.traffic.  Usage: python tools/ncu_traffic.py report.ncu-rep kernel_regex out.json [per_frame_regex]
With per_frame_regex (e.g. k_project) the matched launches are SUMMED per frame: frames = launches matching that regex
(a stage made of several kernels, like k_cull + k_project)."""
import csv
import io
import json
import re
import subprocess
import sys

rep, kern, out = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
ri, wi, ti = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launches = []
for r in rows[2:]:
    if len(r) == len(hdr) and re.search(kern, r[ki]):
        launches.append(dict(read=float(r[ri]) * scale[units[ri]], write=float(r[wi]) * scale[units[wi]],
                             duration=float(r[ti]), duration_unit=units[ti]))
n = len(launches)
if len(sys.argv) > 4:
    n = sum(1 for r in rows[2:] if len(r) == len(hdr) and re.search(sys.argv[4], r[ki])) or n
res = {"kernel": kern, "launches": n, "report": rep.split("/")[-1],
       "dram_bytes_read_per_launch": sum(l["read"] for l in launches) / n,
       "dram_bytes_write_per_launch": sum(l["write"] for l in launches) / n,
       "ncu_duration_per_launch": sum(l["duration"] for l in launches) / n, "ncu_duration_unit": launches[0]["duration_unit"]}
res["dram_bytes_per_launch"] = res["dram_bytes_read_per_launch"] + res["dram_bytes_write_per_launch"]
json.dump(res, open(out, "w"), indent=1)
print(res)
