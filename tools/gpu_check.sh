#!/bin/bash
# One GPU round trip: parity tests, the bench line, and the per-launch device times of one steady-state frame.
# usage (under gpurun): bash tools/gpu_check.sh <tag> [pytest args]
tag=${1:-run}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q ${@:2} > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "stages", {k: round(v, 4) for k, v in d["stages_ms"].items()},
          "V", d["config"]["visible_mean"], "pairs", d["config"]["pairs_mean"], "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("bench failed:", e)
    print(open("gpurun_out/${tag}_bench.err").read()[-2000:])
PY
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
