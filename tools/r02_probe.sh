#!/bin/bash
# Round-2 first GPU round trip: host facts + Vulkan probe, the full-size image parity tests, fp32 / unorm8 bench lines.
tag=${1:-r02a}
mkdir -p gpurun_out
{
  echo "== host"; nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
  echo "== vulkan probe"
  ls /usr/lib/x86_64-linux-gnu 2>/dev/null | grep -i -E "vulkan|lvp|llvmpipe|libGL|libEGL" || echo "no libvulkan / lavapipe / GL libraries under /usr/lib/x86_64-linux-gnu"
  ldconfig -p | grep -i vulkan || echo "ldconfig: no vulkan"
  which vulkaninfo glslangValidator glslc slangc Xvfb 2>&1 || echo "no vulkaninfo / glslangValidator / glslc / slangc / Xvfb on PATH"
  ls /usr/share/vulkan /etc/vulkan /usr/share/glvnd 2>&1
  find / -xdev \( -name "libvulkan*" -o -name "*lvp_icd*" -o -name "nvidia_icd*.json" \) 2>/dev/null | head -20
  echo "== end"
} > gpurun_out/${tag}_host_vulkan_probe.txt 2>&1
cat gpurun_out/${tag}_host_vulkan_probe.txt
timeout 1500 python -m pytest tests/test_gpu_image_parity.py -q --durations=10 > gpurun_out/${tag}_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/${tag}_parity.log
tail -40 gpurun_out/${tag}_parity.log
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
