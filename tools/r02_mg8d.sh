#!/bin/bash
# 8 GPUs: C5 in screen bands, replicated cull with tile boxes, band edges re-cut from measured band times
tag=${1:-r02mg8d}; n=${2:-8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --config c5 --steps 20 --warmup 4 > gpurun_out/${tag}_c5.json 2> gpurun_out/${tag}_c5.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_c5.json").read().strip().splitlines()[-1])
    print("c5: value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["stages_ms"].items()}, d.get("slowest_band_stages_ms_total"), d["config"].get("band_edges"))
    for b in d.get("bands") or []: print("    ", b)
except Exception as e:
    print("c5: failed:", e); print(open("gpurun_out/${tag}_c5.err").read()[-2500:])
PY
