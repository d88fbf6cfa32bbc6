#!/bin/bash
# same-box A/B of two builds of the library: tools/_ab/libsfd2_b200_old.so (previous commit) against the in-tree one
tag=${1:-ab}
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
AB_LIB=tools/_ab/libsfd2_b200_old.so timeout 100 python tools/ab_bench.py "old:mixed:" "old:exact:" >> gpurun_out/${tag}_ab.log 2>&1
timeout 100 python tools/ab_bench.py "new:mixed:" "new:exact:" "new:fast:" >> gpurun_out/${tag}_ab.log 2>&1
grep -A1 "^##" gpurun_out/${tag}_ab.log
