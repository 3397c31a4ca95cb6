#!/bin/bash
tag=${1:-r02g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_configs.py -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
for v in main x_COALPOS; do
  lib=""
  [ "$v" != "main" ] && lib="$PWD/vkgs_b200/lib/libvkgsb_${v}.so"
  VKGSB_LIB=$lib timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_${v}.json 2> gpurun_out/${tag}_${v}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${v}.json"))
    print("${v}: fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "roof", round(d["roofline"]["frac"], 3), "u8", round(d.get("value_unorm8", 0), 1))
except Exception as e:
    print("${v}: bench failed:", e); print(open("gpurun_out/${tag}_${v}.err").read()[-1500:])
PY
done
