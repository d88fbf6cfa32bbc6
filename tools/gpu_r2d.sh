#!/bin/bash
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 300 python tools/match_bench.py exact > gpurun_out/${tag}_match_exact.txt 2>&1; cat gpurun_out/${tag}_match_exact.txt
timeout 300 python tools/match_bench.py fast > gpurun_out/${tag}_match_fast.txt 2>&1; head -3 gpurun_out/${tag}_match_fast.txt
for dbg in 1 2 4 6; do
  echo "--- SFD2_TM_DEBUG=$dbg"
  SFD2_TM_DEBUG=$dbg timeout 300 python tools/match_bench.py exact > gpurun_out/${tag}_match_dbg${dbg}.txt 2>&1; cat gpurun_out/${tag}_match_dbg${dbg}.txt
done
timeout 600 python -m pytest tests -m gpu -x -q -k "match or grouped or hloc_layout or localizer or ratio or pair_pipeline" 2>&1 | tail -3
echo done
