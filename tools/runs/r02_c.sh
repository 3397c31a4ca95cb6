#!/bin/bash
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_overlay.py tests/test_gpu_configs.py tests/test_gpu_image_parity.py -x -q -k "not c5" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
for b in fp32 unorm8; do
  timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --blend $b > gpurun_out/${tag}_bench_$b.json 2> gpurun_out/${tag}_bench_$b.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_$b.json"))
    print("$b fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "stages", {k: round(v, 4) for k, v in d["stages_ms"].items()}, "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("bench failed:", e); print(open("gpurun_out/${tag}_bench_$b.err").read()[-2000:])
PY
done
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_cull|k_project" -o gpurun_out/${tag}_project python tools/profile_frame.py --frames 1 > gpurun_out/${tag}_ncu.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_blend" -o gpurun_out/${tag}_blend8 python tools/profile_frame.py --frames 1 --blend unorm8 > gpurun_out/${tag}_ncu2.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log gpurun_out/${tag}_ncu2.log
