"""Small driver for ncu: a few matcher calls (single 4096^2 pair: mutual / rows only / ratio; one-to-many)."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from sfd2_b200.matchers import match_dev, match_one_to_many
from sfd2_b200.synth import synth_descriptors
prec = sys.argv[1] if len(sys.argv) > 1 else "exact"
d0, d1 = synth_descriptors(0, 4096, 4096)
a, b = torch.from_numpy(d0).cuda(), torch.from_numpy(d1).cuda()
for kw in ({}, {}, {"mutual": False}, {"ratio_th": 0.8}):
    match_dev(a, b, precision=prec, **kw)
rng = np.random.RandomState(100)
dbs = rng.randn(50 * 2000, 128).astype(np.float32)
dbs /= np.linalg.norm(dbs, axis=1, keepdims=True)
match_one_to_many(a, torch.from_numpy(dbs).cuda(), np.arange(51, dtype=np.int32) * 2000, precision=prec)
torch.cuda.synchronize()
print("ok")
