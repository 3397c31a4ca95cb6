#!/bin/bash
# last check of the committed code: smoke(), the parity tests without the 50 M scene, stage times of C3 / C4 on one GPU
tag=${1:-r02finb}
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
tail -4 gpurun_out/${tag}_smoke.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_group.py tests/test_gpu_public_surface.py tests/test_gpu_external.py -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python tools/config_times.py c3 c4 > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err
cat gpurun_out/${tag}_configs.jsonl; tail -2 gpurun_out/${tag}_configs.err
