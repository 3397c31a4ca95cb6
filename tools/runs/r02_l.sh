#!/bin/bash
tag=${1:-r02l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_full_size.py -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 900 python tools/config_times.py c5 > gpurun_out/${tag}_c5.jsonl 2> gpurun_out/${tag}_c5.err; cat gpurun_out/${tag}_c5.jsonl; tail -3 gpurun_out/${tag}_c5.err
timeout 600 python bench.py --config c4 --steps 360 --warmup 5 > gpurun_out/${tag}_c4_360.json 2> gpurun_out/${tag}_c4_360.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_c4_360.json").read().strip().splitlines()[-1])
print("c4 360 views on 1 GPU:", round(d["value"], 1), "views/s", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["stages_ms"].items()})
PY
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_blend" -o gpurun_out/${tag}_blend python tools/profile_frame.py --frames 1 > gpurun_out/${tag}_ncu.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_blend" -o gpurun_out/${tag}_blend8 python tools/profile_frame.py --frames 1 --blend unorm8 > gpurun_out/${tag}_ncu2.log 2>&1
