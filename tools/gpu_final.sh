#!/bin/bash
# Round-end evidence run (one gpurun call): tests, smoke, the bench (default = mixed) and the other precision modes, the CPU
# reference arm, the ncu launch list of the default bench command, ncu --set full captures of the extract path and of the
# matcher, and SASS listings of the tensor-core kernels.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-final}
KREGEX='regex:conv|nms|select|sample|heat|match|preprocess|desc'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_mixed.json 2> gpurun_out/${tag}_bench_mixed.err
python tools/show_bench.py gpurun_out/${tag}_bench_mixed.json > gpurun_out/${tag}_show.txt 2>&1; head -30 gpurun_out/${tag}_show.txt | cut -c1-600
for prec in exact fast; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-other-modes --precision $prec > gpurun_out/${tag}_bench_${prec}.json 2> gpurun_out/${tag}_bench_${prec}.err
  python tools/show_bench.py gpurun_out/${tag}_bench_${prec}.json 2>&1 | head -1 | cut -c1-300
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
cut -c1-400 gpurun_out/${tag}_bench_reference.json
timeout 300 python tools/match_bench.py exact > gpurun_out/${tag}_match_bench_exact.txt 2>&1
timeout 300 python tools/single_call.py mixed > gpurun_out/${tag}_single_call.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-other-modes --pairs 16 > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k "$KREGEX" --launch-skip 23 -c 23 -o gpurun_out/${tag}_prof_mixed python tools/profile_extract.py mixed 2 > gpurun_out/${tag}_prof_mixed.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tc_match_kernel|match_prep_kernel" -c 10 -o gpurun_out/${tag}_prof_match python tools/profile_match.py exact > gpurun_out/${tag}_prof_match.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
echo done
