#!/bin/bash
# round-2 first GPU pass: matcher tests first (new kernel), then the whole GPU suite, then a short bench
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "match or grouped or hloc_layout or localizer or ratio" > gpurun_out/${tag}_pytest_match.log 2>&1
echo "pytest-match exit $?" >> gpurun_out/${tag}_pytest_match.log
tail -15 gpurun_out/${tag}_pytest_match.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-other-modes > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
python tools/show_bench.py gpurun_out/${tag}_bench.json 2>/dev/null | head -30
echo done
