"""Per-layer times of one 1600x1200 extraction (library's per-launch events), for A/B runs with env knobs."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from sfd2_b200 import Extractor
from sfd2_b200.synth import synth_image_u8
prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
ex = Extractor(os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz"), precision=prec, topk=4096)
imgs = torch.from_numpy(np.stack([synth_image_u8(s, 1200, 1600) for s in range(2)])).cuda()
for i in range(3): ex(imgs[i % 2:i % 2 + 1])
ctx = ex.model.ctx
ctx.profile(True); ctx.profile_read()
for i in range(8): ex(imgs[i % 2:i % 2 + 1])
pr = ctx.profile_read(); ctx.profile(False)
print({k.split(":")[-1]: round(v[1] / v[0] * 1e3, 1) for k, v in pr.items()}, "total", round(sum(v[1] for v in pr.values()) / 8, 3), "ms", {k: v for k, v in os.environ.items() if k.startswith("SFD2_")})
