#!/bin/bash
# A/B of library variants on the GPU box: bench line (stage times) per variant.  usage: tools/ab_bench.sh <tag> [variant ...]
# "main" = the product library; other names = vkgs_b200/lib/libvkgsb_<name>.so (tools/build_variant.sh).
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=""
  [ "$v" != "main" ] && lib="$PWD/vkgs_b200/lib/libvkgsb_${v}.so"
  VKGSB_LIB=$lib timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/${tag}_${v}.json 2> gpurun_out/${tag}_${v}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${v}.json"))
    print("${v}: fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("${v}: bench failed:", e); print(open("gpurun_out/${tag}_${v}.err").read()[-1500:])
PY
done
