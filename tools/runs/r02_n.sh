#!/bin/bash
tag=${1:-r02n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "not c5_50m" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_main.json 2> gpurun_out/${tag}_main.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_main.json"))
    print("main: fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "roof", round(d["roofline"]["frac"], 3), "u8", round(d.get("value_unorm8", 0), 1), {k: round(v, 4) for k, v in d["stages_ms_unorm8"].items()}, d["blend"], d["blend_unorm8"])
except Exception as e:
    print("main: bench failed:", e); print(open("gpurun_out/${tag}_main.err").read()[-800:])
PY
