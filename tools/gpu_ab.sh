#!/bin/bash
# A/B run of the library toggles (see tools/ab_bench.py); results in gpurun_out/<tag>_ab.log
tag=${1:-ab}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 400 python tools/ab_bench.py "warm:mixed:" "base:mixed:" "base:exact:" "base:fast:" > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
