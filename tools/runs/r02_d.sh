#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_overlay.py tests/test_gpu_image_parity.py -x -q -k "not c5" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
for b in unorm8; do
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
bash tools/ab_bench.sh $tag main x_NOPOS x_NOHIST x_NOSTORE x_NOLOAD x_NOMATH
