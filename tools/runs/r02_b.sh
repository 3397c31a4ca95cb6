#!/bin/bash
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_sort.py tests/test_gpu_overlay.py -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
bash tools/ab_bench.sh $tag main b4r2 b3r2 b2r4
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_frame.py --frames 2 > gpurun_out/${tag}_pf.log 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/${tag}_launches.csv") if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
n = len(rows) // 2
tot = 0
for r in rows[-n:]:
    tot += float(r[vi])
    print(f"{float(r[vi])/1e3:9.1f} us {r[gi]:>14}  {r[ki][:60]}")
print(f"{tot/1e3:9.1f} us total")
PY
