import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from sfd2_b200 import Matcher, matcher_confs
from sfd2_b200.matchers import match_dev
g = np.load(os.path.join(REPO, "tests/golden/c1_640x480.npz"))
d0, d1 = g["desc"].astype(np.float32), g["desc_b"].astype(np.float32)
gm = np.load(os.path.join(REPO, "tests/golden/match_cases.npz"))
ref = gm["c1_itloc_nnr_m0"]
for rep in range(3):
    o2 = Matcher(matcher_confs["NNR"], precision="exact")({"descriptors0": d0.astype(np.float64), "descriptors1": d1.astype(np.float64)})
    m = o2["matches0"]
    print("rep", rep, "mismatch", int((m != ref).sum()), "ours matched", int((m >= 0).sum()), "ref matched", int((ref >= 0).sum()))
sim = d0.astype(np.float64) @ d1.astype(np.float64).T
def ratio_ok(s0, s1, r=0.9):
    return np.sqrt(2 - 2 * s0) / (np.sqrt(2 - 2 * s1) + 1e-8) <= r
srt = np.sort(sim, axis=1); rb, rs = srt[:, -1], srt[:, -2]
srt = np.sort(sim, axis=0); cb, cs = srt[-1], srt[-2]
nn12, nn21 = sim.argmax(1), sim.argmax(0)
mutual = nn21[nn12] == np.arange(len(d0))
full = np.where(mutual & ratio_ok(rb, rs) & ratio_ok(cb[nn12], cs[nn12]), nn12, -1)
rows_only = np.where(mutual & ratio_ok(rb, rs), nn12, -1)
cols_only = np.where(mutual & ratio_ok(cb[nn12], cs[nn12]), nn12, -1)
print("fp64 full vs ref", int((full != ref).sum()), "| ours vs full", int((m != full).sum()), "ours vs rows_only", int((m != rows_only).sum()),
      "ours vs cols_only", int((m != cols_only).sum()), "ours vs mutual-only", int((m != np.where(mutual, nn12, -1)).sum()))
bad = np.nonzero(m != ref)[0][:10]
for i in bad:
    print(i, "ours", m[i], "ref", ref[i], "row best/sec", rb[i], rs[i], "col best/sec", cb[nn12[i]], cs[nn12[i]], "col tile", nn12[i] // 128, "row tile", i // 128,
          "row2 col", int(np.argsort(sim[i])[-2]), "col2 row", int(np.argsort(sim[:, nn12[i]])[-2]))
a, b = torch.from_numpy(d0).cuda(), torch.from_numpy(d1).cuda()
m2, _ = match_dev(a, b, ratio_th=0.9, ratio_mode=1, precision="exact")
print("match_dev device path mismatch vs ref", int((m2.cpu().numpy() != ref).sum()))
print("---- modes")
def hloc_ok(s0, s1, r):
    return 2 * (1 - s0) <= r * r * 2 * (1 - s1)
for mode, fn in ((0, hloc_ok), (1, ratio_ok)):
    for r in (0.8, 0.9, 0.95):
        mm, _ = match_dev(a, b, ratio_th=r, ratio_mode=mode, precision="exact")
        mm = mm.cpu().numpy()
        full = np.where(mutual & fn(rb, rs, r) & fn(cb[nn12], cs[nn12], r), nn12, -1)
        ro = np.where(mutual & fn(rb, rs, r), nn12, -1)
        co = np.where(mutual & fn(cb[nn12], cs[nn12], r), nn12, -1)
        mo = np.where(mutual, nn12, -1)
        print("mode", mode, "r", r, "vs full", int((mm != full).sum()), "rows_only", int((mm != ro).sum()), "cols_only", int((mm != co).sum()), "mutual-only", int((mm != mo).sum()),
              "| nomutual:", end=" ")
        mn, _ = match_dev(a, b, mutual=False, ratio_th=r, ratio_mode=mode, precision="exact")
        print(int((mn.cpu().numpy() != np.where(fn(rb, rs, r), nn12, -1)).sum()))
