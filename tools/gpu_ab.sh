#!/bin/bash
# A/B run of the library toggles (see tools/ab_bench.py); results in gpurun_out/<tag>_ab.log
tag=${1:-ab}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
B="SFD2_TC_PREFETCH=1,SFD2_TC_BSTAGES=12,SFD2_FUSE_STA=1,SFD2_STREAMS=2,SFD2_TC_MULTICAST=1"
timeout 900 python tools/ab_bench.py \
  "base:mixed:$B" \
  "nopf:mixed:${B/SFD2_TC_PREFETCH=1/SFD2_TC_PREFETCH=0}" \
  "bs8:mixed:${B/SFD2_TC_BSTAGES=12/SFD2_TC_BSTAGES=8}" \
  "nosta:mixed:${B/SFD2_FUSE_STA=1/SFD2_FUSE_STA=0}" \
  "s3:mixed:${B/SFD2_STREAMS=2/SFD2_STREAMS=3}" \
  "s4:mixed:${B/SFD2_STREAMS=2/SFD2_STREAMS=4}" \
  "nomc:mixed:${B/SFD2_TC_MULTICAST=1/SFD2_TC_MULTICAST=0}" \
  "base:exact:$B" "nopf:exact:${B/SFD2_TC_PREFETCH=1/SFD2_TC_PREFETCH=0}" \
  "base:fast:$B" "nopf:fast:${B/SFD2_TC_PREFETCH=1/SFD2_TC_PREFETCH=0}" \
  > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
