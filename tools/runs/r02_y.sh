#!/bin/bash
# node-level expansion of sparse nodes in k_project: parity, the headline bench, C5 band stage times on one GPU
tag=${1:-r02y}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 12 --no-cpu-baseline > gpurun_out/${tag}_main.json 2> gpurun_out/${tag}_main.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_main.json"))
    print("main: fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "roof", round(d["roofline"]["frac"], 3), "u8", round(d.get("value_unorm8", 0), 1))
except Exception as e:
    print("main: bench failed:", e); print(open("gpurun_out/${tag}_main.err").read()[-1500:])
PY
timeout 600 python tools/band_stages.py > gpurun_out/${tag}_bands.jsonl 2> gpurun_out/${tag}_bands.err
cat gpurun_out/${tag}_bands.jsonl; tail -3 gpurun_out/${tag}_bands.err
