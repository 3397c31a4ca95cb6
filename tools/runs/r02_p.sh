#!/bin/bash
tag=${1:-r02p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_external.py tests/test_gpu_shared.py tests/test_gpu_group.py tests/test_gpu_parity.py -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -12 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --config c4 --steps 360 --warmup 8 > gpurun_out/${tag}_c4_360.json 2> gpurun_out/${tag}_c4_360.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_c4_360.json").read().strip().splitlines()[-1])
    print("c4 360 views on 1 GPU:", round(d["value"], 1), "views/s", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["stages_ms"].items()})
except Exception as e:
    print("c4 failed", e); print(open("gpurun_out/${tag}_c4_360.err").read()[-1500:])
PY
timeout 900 python bench.py --config c5 --steps 20 --warmup 4 > gpurun_out/${tag}_c5.json 2> gpurun_out/${tag}_c5.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_c5.json").read().strip().splitlines()[-1])
    print("c5 on 1 GPU:", round(d["value"], 1), "fps", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["stages_ms"].items()})
except Exception as e:
    print("c5 failed", e); print(open("gpurun_out/${tag}_c5.err").read()[-1500:])
PY
