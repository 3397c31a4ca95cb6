#!/bin/bash
tag=${1:-r02h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shared.py -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for v in main b4r2 x_COALPOS; do
  lib=""
  [ "$v" != "main" ] && lib="$PWD/vkgs_b200/lib/libvkgsb_${v}.so"
  VKGSB_LIB=$lib timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_${v}.json 2> gpurun_out/${tag}_${v}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${v}.json"))
    print("${v}: fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "roof", round(d["roofline"]["frac"], 3), "u8", round(d.get("value_unorm8", 0), 1))
except Exception as e:
    print("${v}: bench failed:", e); print(open("gpurun_out/${tag}_${v}.err").read()[-800:])
PY
done
for v in main b4r2; do
lib=""
[ "$v" != "main" ] && lib="$PWD/vkgs_b200/lib/libvkgsb_${v}.so"
VKGSB_LIB=$lib timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/${tag}_launches_$v.csv python tools/profile_frame.py --frames 2 > gpurun_out/${tag}_pf.log 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/${tag}_launches_$v.csv") if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
n = len(rows) // 2
tot = 0
for r in rows[-n:]:
    tot += float(r[vi])
    print(f"{float(r[vi])/1e3:9.1f} us {r[gi]:>14}  {r[ki][:60]}")
print(f"{tot/1e3:9.1f} us total ($v, cache-control none: warm L2)")
PY
done
