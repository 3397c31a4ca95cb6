#!/bin/bash
tag=${1:-r02mg2b}; n=${2:-2}; n5=${3:-50000000}
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${name}.json").read().strip().splitlines()[-1])
    print("${name}: value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "slowest", d.get("slowest_band_stages_ms_total"), d["config"].get("band_edges"))
    for b in d.get("bands") or []: print("    ", b)
except Exception as e:
    print("${name}: failed:", e); print(open("gpurun_out/${tag}_${name}.err").read()[-2500:])
PY
}
run c5_shared --config c5 --steps 20 --warmup 3 --band-cull shared --n-splats $n5
run c5_repl --config c5 --steps 20 --warmup 3 --band-cull replicated --n-splats $n5
