#!/bin/bash
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 300 python tools/match_bench.py exact > gpurun_out/${tag}_match_exact.txt 2>&1; cat gpurun_out/${tag}_match_exact.txt
timeout 300 python tools/match_bench.py fast > gpurun_out/${tag}_match_fast.txt 2>&1; cat gpurun_out/${tag}_match_fast.txt
timeout 600 python -m pytest tests -m gpu -q -k "match or grouped or hloc_layout or localizer or ratio or pair_pipeline" > gpurun_out/${tag}_pytest_match.log 2>&1
grep -E "^E  |passed|failed|Error" gpurun_out/${tag}_pytest_match.log | head -40
for i in 1 2 3; do timeout 300 python -m pytest tests -m gpu -q -k "ratio" 2>&1 | tail -1; done
echo done
