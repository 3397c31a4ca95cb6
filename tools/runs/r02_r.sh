#!/bin/bash
# quick GPU check of a change: the parity tests that do not need the 50 M scene, then the headline bench
tag=${1:-r02r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_group.py tests/test_gpu_public_surface.py -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 12 --no-cpu-baseline > gpurun_out/${tag}_main.json 2> gpurun_out/${tag}_main.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_main.json"))
    print("main: fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "roof", round(d["roofline"]["frac"], 3), "u8", round(d.get("value_unorm8", 0), 1))
except Exception as e:
    print("main: bench failed:", e); print(open("gpurun_out/${tag}_main.err").read()[-1500:])
PY
