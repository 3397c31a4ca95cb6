#!/bin/bash
# parity tests (no 50 M scene) + ncu --set full of the projection stage's kernels on the bench frame
tag=${1:-r02s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_group.py tests/test_gpu_public_surface.py -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/${tag}_launches_warm.csv python tools/profile_frame.py --frames 2 > gpurun_out/${tag}_pf.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --cache-control none --import-source on -k regex:"k_project|k_cull" -o gpurun_out/${tag}_proj python tools/profile_frame.py --frames 1 > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
grep -v "^==" gpurun_out/${tag}_launches_warm.csv | cut -d, -f5,14-16 | tail -24
