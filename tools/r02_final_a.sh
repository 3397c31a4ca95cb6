#!/bin/bash
# Round-2 final check on one B200: the whole GPU test suite, the bench lines (c2 default, c5 whole frame), the launch list
# of one steady-state frame and an `ncu --set full` capture of it (numbers printed under ncu are not bench values).
tag=${1:-r02fin}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${tag}_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_full.log
tail -14 gpurun_out/${tag}_pytest_full.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 400 python bench.py --config c5 --steps 20 --warmup 4 > gpurun_out/${tag}_c5_1gpu.json 2> gpurun_out/${tag}_c5_1gpu.err
python - <<PY
import json
for f in ("bench", "c5_1gpu"):
    try:
        d = json.loads(open("gpurun_out/${tag}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("value", 0), 1), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "roof", round(d.get("roofline", {}).get("frac", 0), 3), "u8", round(d.get("value_unorm8", 0), 1), "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "failed:", e); print(open("gpurun_out/${tag}_%s.err" % f).read()[-1500:])
PY
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/${tag}_launches_warm.csv python tools/profile_frame.py --frames 2 > gpurun_out/${tag}_pf.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --cache-control none --import-source on -o gpurun_out/${tag}_frame python tools/profile_frame.py --frames 1 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
