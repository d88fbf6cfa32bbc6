#!/bin/bash
# A/B run of the library toggles (see tools/ab_bench.py); results in gpurun_out/<tag>_ab.log
tag=${1:-ab}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
B="SFD2_TC_PREFETCH=0,SFD2_TC_BSTAGES=12,SFD2_FUSE_STA=1,SFD2_STREAMS=2,SFD2_TC_MULTICAST=1"
timeout 900 python tools/ab_bench.py "base:mixed:$B" "nosta:mixed:${B/SFD2_FUSE_STA=1/SFD2_FUSE_STA=0}" "base:exact:$B" "base:fast:$B" > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
python - <<'PY' 2>&1 | tee gpurun_out/${tag}_wbw.log
import torch
x = torch.empty(491_520_000 // 4, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in [("zero_ (pure write 491 MB)", lambda: x.zero_()), ("copy_ (read+write 2x491 MB)", lambda: y.copy_(x)),
                 ("sum (pure read 491 MB)", lambda: x.sum())]:
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms*1000:.1f} us -> {491.52e6 / ms / 1e6:.0f} GB/s per 491 MB")
PY
