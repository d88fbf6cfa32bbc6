#!/bin/bash
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tc_match_kernel|match_prep_kernel" -c 10 -o gpurun_out/${tag}_prof_match python tools/profile_match.py exact > gpurun_out/${tag}_prof_match.log 2>&1
tail -3 gpurun_out/${tag}_prof_match.log
ls -la gpurun_out/${tag}_*.ncu-rep
echo done
