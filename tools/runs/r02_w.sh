#!/bin/bash
# tile-shuffle experiment on the bench orbit + C5 (50 M splats) whole frame on one GPU
tag=${1:-r02w}
mkdir -p gpurun_out
run() {
  v=$1; shift
  timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/${tag}_${v}.json 2> gpurun_out/${tag}_${v}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${v}.json").read().strip().splitlines()[-1])
    print("${v}: value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("value", 0), 1), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "roof", round(d.get("roofline", {}).get("frac", 0), 3), "u8", round(d.get("value_unorm8", 0), 1))
except Exception as e:
    print("${v}: bench failed:", e); print(open("gpurun_out/${tag}_${v}.err").read()[-1500:])
PY
}
run main --steps 100 --warmup 12
VKGSB_SPATIAL_SHUFFLE=1 run shuffle --steps 100 --warmup 12
run c5 --config c5 --steps 20 --warmup 4
