#!/bin/bash
# Round-end evidence run (one gpurun call): tests, smoke, the bench in every mode, the CPU reference arm, the ncu launch
# list of the default bench command and one ncu --set full capture.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-final}
KREGEX='regex:conv|nms|select|sample|heat|match|split|norm|sta_'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_mixed.json 2> gpurun_out/${tag}_bench_mixed.err
python tools/show_bench.py gpurun_out/${tag}_bench_mixed.json > gpurun_out/${tag}_show.txt 2>&1; head -5 gpurun_out/${tag}_show.txt
for prec in exact fast; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-other-modes --precision $prec > gpurun_out/${tag}_bench_${prec}.json 2> gpurun_out/${tag}_bench_${prec}.err
  python tools/show_bench.py gpurun_out/${tag}_bench_${prec}.json > gpurun_out/${tag}_show_${prec}.txt 2>&1; head -1 gpurun_out/${tag}_show_${prec}.txt
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
cut -c1-300 gpurun_out/${tag}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 500 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-other-modes > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 800 ncu --set full --clock-control none -k "$KREGEX" --launch-skip 23 -c 23 -o gpurun_out/${tag}_prof_mixed python tools/profile_extract.py mixed 2 > gpurun_out/${tag}_prof_mixed.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:tc_match|split_rows2|match_finish" -c 6 -o gpurun_out/${tag}_prof_match python tools/profile_extract.py mixed 1 > gpurun_out/${tag}_prof_match.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
echo done
