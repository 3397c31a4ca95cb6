#!/bin/bash
tag=${1:-r02j}
mkdir -p gpurun_out
for v in main bl_f4 bl_f2 bl_u2; do
  lib=""
  [ "$v" != "main" ] && lib="$PWD/vkgs_b200/lib/libvkgsb_${v}.so"
  VKGSB_LIB=$lib timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_${v}.json 2> gpurun_out/${tag}_${v}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${v}.json"))
    print("${v}: fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "u8", round(d.get("value_unorm8", 0), 1), "u8 blend", round(d["stages_ms_unorm8"]["blend"], 4), "p10/50/90", [round(d["frame_ms"][k], 3) for k in ("p10", "p50", "p90")])
except Exception as e:
    print("${v}: bench failed:", e); print(open("gpurun_out/${tag}_${v}.err").read()[-800:])
PY
done
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_bin|k_sort" -o gpurun_out/${tag}_binsort python tools/profile_frame.py --frames 1 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
