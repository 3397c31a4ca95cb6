#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_overlay.py tests/test_gpu_sort.py tests/test_gpu_full_size.py -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
show() {
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_$1.json"))
    print("$1: fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "roof", round(d["roofline"]["frac"], 3),
          "| u8 fps", round(d.get("value_unorm8", 0), 1), {k: round(v, 4) for k, v in d.get("stages_ms_unorm8", {}).items()}, "retries", d.get("blend_unorm8", {}).get("retries_per_frame"),
          "frag", d.get("blend", {}).get("fragments_per_frame"), d.get("blend_unorm8", {}).get("fragments_per_frame"))
except Exception as e:
    print("$1: bench failed:", e); print(open("gpurun_out/${tag}_$1.err").read()[-1500:])
PY
}
for pin in 72 0 48 96; do
  VKGSB_L2_PIN_MB=$pin timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_pin$pin.json 2> gpurun_out/${tag}_pin$pin.err
  show pin$pin
done
for cut in 4 6 7; do
  VKGSB_UNORM8_CUT_EXP=$cut timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_cut$cut.json 2> gpurun_out/${tag}_cut$cut.err
  show cut$cut
done
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
