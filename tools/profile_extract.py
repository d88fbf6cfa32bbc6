"""Small driver for ncu: a few 1600x1200 extractions + matches through the public API."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from sfd2_b200 import Extractor
from sfd2_b200.matchers import match_dev
from sfd2_b200.synth import synth_image_u8, synth_descriptors

prec = sys.argv[1] if len(sys.argv) > 1 else "exact"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ex = Extractor(os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz"), precision=prec, topk=4096)
imgs = torch.from_numpy(np.stack([synth_image_u8(s, 1200, 1600) for s in range(2)])).cuda()
for i in range(n):
    out = ex(imgs[i % 2: i % 2 + 1])
torch.cuda.synchronize()
d0, d1 = synth_descriptors(0, 4096, 4096)
a, b = torch.from_numpy(d0).cuda(), torch.from_numpy(d1).cuda()
for i in range(2):
    match_dev(a, b, precision=prec)
torch.cuda.synchronize()
print("counts", out["counts"].tolist())
