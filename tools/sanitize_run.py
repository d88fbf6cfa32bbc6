"""Small driver for compute-sanitizer: every kernel of the hot path once, on small inputs (all precision modes, banded host
upload, grouped matcher with ids / ratio).  compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from sfd2_b200 import Extractor
from sfd2_b200.matchers import match_dev, match_one_to_many
from sfd2_b200.synth import synth_image_u8, synth_descriptors
W = os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz")
for prec in ("mixed", "exact", "fast"):
    ex = Extractor(W, precision=prec, topk=600)
    for (h, w) in ((200, 264), (97, 141)):
        img = torch.from_numpy(np.stack([synth_image_u8(s, h, w) for s in range(2)])).cuda()
        out = ex(img)
        print(prec, h, w, out["counts"].tolist())
    big = torch.from_numpy(synth_image_u8(7, 1200, 1600)[None]).pin_memory()      # >= 4 MB: banded upload path
    print(prec, "host banded", ex.extract_host(big)["counts"].tolist())
d0, d1 = synth_descriptors(0, 700, 900)
a, b = torch.from_numpy(d0).cuda(), torch.from_numpy(d1).cuda()
for kw in ({}, {"mutual": False}, {"ratio_th": 0.8}):
    m, s = match_dev(a, b, precision="exact", **kw)[:2]
    print("match", kw, int((m >= 0).sum()))
ids = torch.from_numpy((np.arange(1800) % 3 - 1).astype(np.int64)).cuda()
db = torch.from_numpy(synth_descriptors(3, 1800, 8)[0]).cuda()
m, s = match_one_to_many(a, db, [0, 900, 1800], db_ids=ids, precision="exact")
print("o2m", (m >= 0).sum(dim=1).tolist())
torch.cuda.synchronize()
print("done")
