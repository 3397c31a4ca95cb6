#!/bin/bash
# Round-2 profiles of the final frame on one B200 (numbers printed under ncu are not bench values).
tag=${1:-r02fin}
mkdir -p gpurun_out
for cc in all none; do
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control $cc --csv --log-file gpurun_out/${tag}_launches_$cc.csv python tools/profile_frame.py --frames 2 > gpurun_out/${tag}_pf.log 2>&1
done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/${tag}_launches_unorm8.csv python tools/profile_frame.py --frames 2 --blend unorm8 > gpurun_out/${tag}_pf.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/${tag}_frame python tools/profile_frame.py --frames 1 > gpurun_out/${tag}_ncu.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_blend -o gpurun_out/${tag}_blend8 python tools/profile_frame.py --frames 1 --blend unorm8 > gpurun_out/${tag}_ncu2.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log gpurun_out/${tag}_ncu2.log
