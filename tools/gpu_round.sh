#!/bin/bash
# One gpurun call: parity tests, then the bench in the three tcgen05 modes, then an ncu launch list.
# usage: tools/gpu_round.sh [tag]
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for prec in exact mixed fast; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --precision $prec > gpurun_out/${tag}_bench_${prec}.json 2> gpurun_out/${tag}_bench_${prec}.err
  python tools/show_bench.py gpurun_out/${tag}_bench_${prec}.json 2>/dev/null | head -40
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --precision mixed > gpurun_out/${tag}_ncu_bench.log 2>&1
echo done
