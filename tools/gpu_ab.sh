#!/bin/bash
# same-box A/B of two builds of the library: tools/_ab/libsfd2_b200_old.so (previous commit) against the in-tree one
tag=${1:-ab}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
for rep in 1; do
  AB_LIB=tools/_ab/libsfd2_b200_old.so timeout 300 python tools/ab_bench.py "old:mixed:" "old:exact:" "old:fast:" >> gpurun_out/${tag}_ab.log 2>&1
  timeout 300 python tools/ab_bench.py "new:mixed:" "new:exact:" "new:fast:" >> gpurun_out/${tag}_ab.log 2>&1
done
grep -A1 "^##" gpurun_out/${tag}_ab.log
