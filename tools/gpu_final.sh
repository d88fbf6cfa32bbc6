#!/bin/bash
# Round-end evidence run (one gpurun call): tests, smoke, the bench in every mode, the CPU reference arm, the ncu launch
# list of the default bench command and one ncu --set full capture.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_mixed.json 2> gpurun_out/${tag}_bench_mixed.err
python tools/show_bench.py gpurun_out/${tag}_bench_mixed.json | head -5
for prec in exact fast; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-other-modes --precision $prec > gpurun_out/${tag}_bench_${prec}.json 2> gpurun_out/${tag}_bench_${prec}.err
  python tools/show_bench.py gpurun_out/${tag}_bench_${prec}.json | head -1
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
cat gpurun_out/${tag}_bench_reference.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sfd2 -c 500 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-other-modes > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 800 ncu --set full --clock-control none -k regex:sfd2 --launch-skip 23 -c 23 -o gpurun_out/${tag}_prof_mixed python tools/profile_extract.py mixed 2 > gpurun_out/${tag}_prof_mixed.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"tc_match|split_rows2|match_finish" -c 6 -o gpurun_out/${tag}_prof_match python tools/profile_extract.py mixed 1 > gpurun_out/${tag}_prof_match.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
echo done
