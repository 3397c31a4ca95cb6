#!/bin/bash
tag=${1:-r02mg8b}; n=${2:-8}
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${name}.json").read().strip().splitlines()[-1])
    print("${name}: value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("value", 0), 1), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "u8", round(d.get("value_unorm8", 0), 1), "u8 e2e", round(d.get("e2e_unorm8", {}).get("value", 0), 1), d.get("slowest_band_stages_ms_total"), d["config"].get("band_edges"))
    for b in d.get("bands") or []: print("    ", b)
except Exception as e:
    print("${name}: failed:", e); print(open("gpurun_out/${tag}_${name}.err").read()[-2500:])
PY
}
run c2_peer --steps 100 --warmup 12 --no-cpu-baseline --gather peer
run c4_peer --config c4 --steps 45 --warmup 8 --gather peer
run c5_shared --config c5 --steps 20 --warmup 4 --band-cull shared
run c5_repl --config c5 --steps 20 --warmup 4 --band-cull replicated
