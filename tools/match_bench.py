"""Matcher micro-benchmark (device-resident): single 4096^2 pair per call, 32 pairs grouped, one query vs 50 x 2000 db.
    python tools/match_bench.py [precision]      (SFD2_TM_ASLOTS=1|2 selects the resident-A configuration)"""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from sfd2_b200.matchers import match_dev, match_sets_dev, match_one_to_many, _ctx
from sfd2_b200.synth import synth_descriptors

prec = sys.argv[1] if len(sys.argv) > 1 else "exact"
dev = torch.device("cuda", 0)
d0, d1 = synth_descriptors(0, 4096, 4096)
a, b = torch.from_numpy(d0).to(dev), torch.from_numpy(d1).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

ctx = _ctx(0)
for kw, name in (({}, "mutual"), ({"mutual": False}, "rows only"), ({"ratio_th": 0.8}, "ratio 0.8 (two products, top-2)")):
    ms = timed(lambda: match_dev(a, b, precision=prec, **kw), 200)
    ctx.profile(True); ctx.profile_read()
    for _ in range(20):
        match_dev(a, b, precision=prec, **kw)
    pr = ctx.profile_read(); ctx.profile(False)
    print(f"single pair [{name}]: {ms * 1e3:.1f} us/call  " + "  ".join(f"{k} {v[1] / v[0] * 1e3:.1f} us" for k, v in pr.items()), flush=True)
G = 32
sets = []
for g in range(G):
    x0, x1 = synth_descriptors(1000 + g, 4096, 4096)
    sets += [{"data": torch.from_numpy(x0).to(dev)}, {"data": torch.from_numpy(x1).to(dev)}]
ga, gb = list(range(0, 2 * G, 2)), list(range(1, 2 * G, 2))
ms = timed(lambda: match_sets_dev(sets, ga, gb, precision=prec), 10)
ctx.profile(True); ctx.profile_read()
for _ in range(5):
    match_sets_dev(sets, ga, gb, precision=prec)
pr = ctx.profile_read(); ctx.profile(False)
print(f"{G} pairs grouped: {ms * 1e3:.1f} us/call = {ms / G * 1e3:.2f} us/pair = {G * 4.295 / ms:.0f} TFLOP/s  " +
      "  ".join(f"{k} {v[1] / v[0] * 1e3:.1f} us" for k, v in pr.items()), flush=True)
rng = np.random.RandomState(100)
dbs = rng.randn(50 * 2000, 128).astype(np.float32)
dbs /= np.linalg.norm(dbs, axis=1, keepdims=True)
dbt = torch.from_numpy(dbs).to(dev)
offs = np.arange(51, dtype=np.int32) * 2000
ms = timed(lambda: match_one_to_many(a, dbt, offs, precision=prec), 10)
ctx.profile(True); ctx.profile_read()
for _ in range(5):
    match_one_to_many(a, dbt, offs, precision=prec)
pr = ctx.profile_read(); ctx.profile(False)
gf = 2.0 * 4096 * 100000 * 128 / 1e9
print(f"one-to-many 4096 x 50 x 2000: {ms * 1e3:.1f} us/call = {gf / ms:.0f} TFLOP/s  " +
      "  ".join(f"{k} {v[1] / v[0] * 1e3:.1f} us ({gf / (v[1] / v[0]):.0f} TF)" for k, v in pr.items()), flush=True)
