#!/bin/bash
tag=${1:-r02c45}
mkdir -p gpurun_out
for c in c4 c5; do
  timeout 1200 python bench.py --config $c --steps 40 --warmup 5 > gpurun_out/${tag}_$c.json 2> gpurun_out/${tag}_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_$c.json").read().strip().splitlines()[-1])
    print("$c: value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("value", 0), 1), {k: round(v, 4) for k, v in d["stages_ms"].items()}, d["config"])
except Exception as e:
    print("$c: failed:", e); print(open("gpurun_out/${tag}_$c.err").read()[-1500:])
PY
done
