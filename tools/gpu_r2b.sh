#!/bin/bash
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "match or grouped or hloc_layout or localizer or ratio or pair or batched" > gpurun_out/${tag}_pytest_match.log 2>&1
echo "pytest-match exit $?" >> gpurun_out/${tag}_pytest_match.log
tail -8 gpurun_out/${tag}_pytest_match.log
for prec in exact fast; do
  timeout 300 python tools/match_bench.py $prec > gpurun_out/${tag}_match_${prec}.txt 2>&1; cat gpurun_out/${tag}_match_${prec}.txt
done
SFD2_TM_ASLOTS=2 timeout 300 python tools/match_bench.py exact > gpurun_out/${tag}_match_exact_aslots2.txt 2>&1; cat gpurun_out/${tag}_match_exact_aslots2.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-other-modes > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -5 gpurun_out/${tag}_bench.err
python tools/show_bench.py gpurun_out/${tag}_bench.json 2>&1 | head -40
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tc_match|match_prep" --launch-skip 8 -c 4 -o gpurun_out/${tag}_prof_match python tools/profile_extract.py mixed 1 > gpurun_out/${tag}_prof_match.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
echo done
