"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference) on CPU.  Build-container only: /root/reference does not exist on
the GPU box, so the fixtures are committed and this script is how they were made.

    python oracle/make_golden.py            # all fixtures
    python oracle/make_golden.py small c1   # a subset

The reference hard-codes .cuda() (nets/extractor.py:106,205; it_loc/matcher.py:93-94)
and it_loc/matcher.py imports h5py at module scope; both are shimmed below
 - no reference source is modified or copied.
"""
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)
torch.Tensor.cuda = lambda s, *a, **k: s
torch.nn.Module.cuda = lambda s, *a, **k: s
sys.modules.setdefault("h5py", types.ModuleType("h5py"))

from nets.sfd2 import ResSegNetV2                                   # noqa: E402
from nets.extractor import extract_resnet_return, simple_nms, norm_RGB  # noqa: E402
from hloc.matchers.nearest_neighbor import NearestNeighbor          # noqa: E402
from it_loc.matcher import Matcher, confs as itloc_confs           # noqa: E402
from sfd2_b200.synth import synth_image_u8, shifted_twin, synth_descriptors  # noqa: E402

WEIGHTS = "/root/reference/weights/20220810_ressegnetv2_wapv2_ce_sd2mfsf_uspg.pth"
OUT = os.path.join(REPO, "tests", "golden")
torch.manual_seed(0)


def model():
    m = ResSegNetV2(outdim=128, require_stability=True).eval()       # extract_localization.py:214
    m.load_state_dict(torch.load(WEIGHTS, map_location="cpu", weights_only=False)["model"], strict=False)
    return m


def to_input(u8):
    return torch.from_numpy((u8.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)[None].copy())


def canon(out, W):
    """Re-sort the reference output by (score desc, y*W+x asc): its own order among
    exactly tied scores is np.argsort-unstable."""
    kp, sc, de = out["keypoints"], out["scores"], out["descriptors"]
    lin = kp[:, 1].astype(np.int64) * W + kp[:, 0].astype(np.int64)
    o = np.lexsort((lin, -sc))
    return kp[o], sc[o], de[o]


def run_extract(m, u8, K):
    img = to_input(u8)
    out = extract_resnet_return(m, img=img, topK=K, mask=None, conf_th=0.001, scales=[1.0])
    kp, sc, de = canon(out, u8.shape[1])
    # the (K+1)-th candidate's score, to detect a tie across the top-K cut
    allout = extract_resnet_return(m, img=img, topK=-1, mask=None, conf_th=0.001, scales=[1.0])
    s_all = np.sort(allout["scores"])[::-1]
    nxt = s_all[K] if len(s_all) > K else -1.0
    return kp, sc, de, len(s_all), nxt


def fixture_extract(m, name, seed, H, W, K, keep_image, with_maps=False, twin=True):
    u8 = synth_image_u8(seed, H, W)
    kp, sc, de, ncand, nxt = run_extract(m, u8, K)
    d = dict(seed=seed, H=H, W=W, K=K, image_sum=np.int64(u8.astype(np.int64).sum()),
             kp_xy=kp.astype(np.int16), scores=sc.astype(np.float32), desc=de.astype(np.float32),
             ncand=np.int64(ncand), next_score=np.float32(nxt))
    if keep_image:
        d["image_u8"] = u8
    if with_maps:
        with torch.no_grad():
            x = norm_RGB(to_input(u8).squeeze())[None]
            hm, st, dm = m.det(x)
            if hm.shape[2] != H or hm.shape[3] != W:
                hm = torch.nn.functional.interpolate(hm, size=[H, W], mode="bilinear", align_corners=False)
            heat = hm * st
            nms = simple_nms(heat, 4)
        d.update(score_map=hm[0, 0].numpy(), stability=st[0, 0].numpy(), desc_map=dm[0].numpy(),
                 heat=heat[0, 0].numpy(), nms=nms[0, 0].numpy())
    if twin:
        u8b = shifted_twin(u8)
        kpb, scb, deb, _, _ = run_extract(m, u8b, K)
        d.update(kp_xy_b=kpb.astype(np.int16), scores_b=scb.astype(np.float32), desc_b=deb.astype(np.float32))
        # both reference matchers on the pair
        nn = NearestNeighbor({"do_mutual_check": True, "distance_threshold": None}).eval()
        d0 = torch.from_numpy(de.T.copy())[None].float()
        d1 = torch.from_numpy(deb.T.copy())[None].float()
        with torch.no_grad():
            ph = nn({"descriptors0": d0, "descriptors1": d1})
        mt = Matcher(conf=itloc_confs["NNM"]).eval()
        pi = mt({"descriptors0": de, "descriptors1": deb})          # float64, as read from h5
        d.update(hloc_matches0=ph["matches0"][0].numpy().astype(np.int32),
                 hloc_scores0=ph["matching_scores0"][0].numpy(),
                 itloc_matches0=np.asarray(pi["matches0"]).astype(np.int32),
                 itloc_scores0=np.asarray(pi["matching_scores0"]).astype(np.float64))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "kpts", len(sc), "cands", ncand, "score[0]", sc[0] if len(sc) else None,
          "score[-1]", sc[-1] if len(sc) else None, "next", nxt)


def fixture_multiscale(m):
    """Multi-scale extraction (nets/extractor.py:113-125, scale loop): one small image, scales [1.0, 0.75, 1.25]."""
    H, W, K = 128, 160, 120
    u8 = synth_image_u8(11, H, W)
    scales = [1.0, 0.75, 1.25]
    out = extract_resnet_return(m, img=to_input(u8), topK=K, mask=None, conf_th=0.001, scales=scales)
    o = np.lexsort((out["keypoints"][:, 1] * 100000 + out["keypoints"][:, 0], -out["scores"]))
    np.savez_compressed(os.path.join(OUT, "ms_128x160.npz"), image_u8=u8, H=H, W=W, K=K, scales=np.array(scales),
                        kp=out["keypoints"][o].astype(np.float32), scores=out["scores"][o].astype(np.float32),
                        desc=out["descriptors"][o].astype(np.float32))
    print("ms_128x160 kpts", len(out["scores"]), out["scores"][:3])


def fixture_nms():
    """simple_nms on hand-made maps: plateaus, exact ties, all-zero, borders."""
    rng = np.random.RandomState(7)
    cases = {}
    a = rng.rand(64, 80).astype(np.float32)
    cases["rand"] = a
    b = np.round(rng.rand(64, 80) * 8).astype(np.float32) / 8          # many exact ties
    cases["ties"] = b
    c = np.zeros((64, 80), np.float32)
    cases["zero"] = c
    d = np.zeros((64, 80), np.float32); d[10:20, 10:30] = 0.5; d[40:44, 60:80] = 0.25; d[0, 0] = 1; d[63, 79] = 1
    cases["plateau"] = d
    e = (np.indices((40, 56)).sum(0) % 2).astype(np.float32) * 0.3       # checkerboard
    cases["checker"] = e
    f = rng.rand(9, 9).astype(np.float32)                               # smaller than the halo
    cases["tiny"] = f
    g = rng.rand(130, 70).astype(np.float32) ** 8                        # sparse peaks, spans tiles
    cases["peaks"] = g
    out = {}
    for k, v in cases.items():
        out["in_" + k] = v
        out["out_" + k] = simple_nms(torch.from_numpy(v)[None, None], 4)[0, 0].numpy()
    np.savez_compressed(os.path.join(OUT, "nms_cases.npz"), **out)
    print("nms_cases", {k: int((out["out_" + k] > 0).sum()) for k in cases})


def fixture_match():
    """Both reference matchers on the randn stress set, ragged shapes included."""
    out = {}
    nn = NearestNeighbor({"do_mutual_check": True, "distance_threshold": None}).eval()
    nn1 = NearestNeighbor({"do_mutual_check": False, "distance_threshold": None}).eval()
    mt = Matcher(conf=itloc_confs["NNM"]).eval()
    mtr = Matcher(conf=itloc_confs["NNR"]).eval()                      # mutual NN + symmetric Lowe ratio 0.9
    nnr = NearestNeighbor({"do_mutual_check": True, "ratio_threshold": 0.8, "distance_threshold": 0.7}).eval()
    nnr1 = NearestNeighbor({"do_mutual_check": False, "ratio_threshold": 0.9, "distance_threshold": None}).eval()
    for tag, (n, m_) in {"sq": (512, 512), "wide": (300, 1000), "tall": (777, 129), "one": (1, 50), "col": (40, 1)}.items():
        d0, d1 = synth_descriptors(11, n, m_)
        with torch.no_grad():
            ph = nn({"descriptors0": torch.from_numpy(d0.T.copy())[None], "descriptors1": torch.from_numpy(d1.T.copy())[None]})
            po = nn1({"descriptors0": torch.from_numpy(d0.T.copy())[None], "descriptors1": torch.from_numpy(d1.T.copy())[None]})
        out[f"{tag}_d0"] = d0; out[f"{tag}_d1"] = d1
        out[f"{tag}_hloc_m0"] = ph["matches0"][0].numpy().astype(np.int32)
        out[f"{tag}_hloc_s0"] = ph["matching_scores0"][0].numpy()
        out[f"{tag}_hloc_nomutual_m0"] = po["matches0"][0].numpy().astype(np.int32)
        if m_ > 1 and n > 1:    # topk(2) needs two candidates in both directions
            with torch.no_grad():
                t0, t1 = torch.from_numpy(d0.T.copy())[None], torch.from_numpy(d1.T.copy())[None]
                pr = nnr({"descriptors0": t0, "descriptors1": t1})
                pr1 = nnr1({"descriptors0": t0, "descriptors1": t1})
            out[f"{tag}_hloc_ratio_m0"] = pr["matches0"][0].numpy().astype(np.int32)
            out[f"{tag}_hloc_ratio_s0"] = pr["matching_scores0"][0].numpy()
            out[f"{tag}_hloc_ratio_nomutual_m0"] = pr1["matches0"][0].numpy().astype(np.int32)
            out[f"{tag}_hloc_ratio_nomutual_s0"] = pr1["matching_scores0"][0].numpy()
            pir = mtr({"descriptors0": d0.astype(np.float64), "descriptors1": d1.astype(np.float64)})
            out[f"{tag}_itloc_nnr_m0"] = np.asarray(pir["matches0"]).astype(np.int32)
        if n > 1:   # the reference's .squeeze() mis-shapes N==1
            pi = mt({"descriptors0": d0.astype(np.float64), "descriptors1": d1.astype(np.float64)})
            out[f"{tag}_itloc_m0"] = np.asarray(pi["matches0"]).astype(np.int32)
            out[f"{tag}_itloc_s0"] = np.asarray(pi["matching_scores0"]).astype(np.float64).reshape(-1)
    # ratio tests on real SFD2 descriptors (the C1 pair stored in c1_640x480.npz): many near-threshold rows
    c1p = os.path.join(OUT, "c1_640x480.npz")
    if os.path.exists(c1p):
        c1 = np.load(c1p)
        d0, d1 = c1["desc"], c1["desc_b"]
        with torch.no_grad():
            t0, t1 = torch.from_numpy(d0.T.copy())[None], torch.from_numpy(d1.T.copy())[None]
            pr = nnr({"descriptors0": t0, "descriptors1": t1})
            pr1 = nnr1({"descriptors0": t0, "descriptors1": t1})
        pir = mtr({"descriptors0": d0.astype(np.float64), "descriptors1": d1.astype(np.float64)})
        out["c1_hloc_ratio_m0"] = pr["matches0"][0].numpy().astype(np.int32)
        out["c1_hloc_ratio_s0"] = pr["matching_scores0"][0].numpy()
        out["c1_hloc_ratio_nomutual_m0"] = pr1["matches0"][0].numpy().astype(np.int32)
        out["c1_hloc_ratio_nomutual_s0"] = pr1["matching_scores0"][0].numpy()
        out["c1_itloc_nnr_m0"] = np.asarray(pir["matches0"]).astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "match_cases.npz"), **out)
    print("match_cases", {k: v.shape for k, v in out.items() if k.endswith("m0")})


def main(which):
    os.makedirs(OUT, exist_ok=True)
    m = model()
    if not which or "small" in which:
        fixture_extract(m, "small_96x128", 3, 96, 128, 50, keep_image=True, with_maps=True)
    if not which or "odd" in which:      # non-multiple-of-8 size: heat-map is bilinearly resized (:137-138)
        fixture_extract(m, "odd_100x141", 5, 100, 141, 60, keep_image=True, with_maps=True, twin=False)
    if not which or "c1" in which:
        fixture_extract(m, "c1_640x480", 0, 480, 640, 1000, keep_image=True)
    if not which or "c2" in which:
        fixture_extract(m, "c2_1600x1200", 0, 1200, 1600, 4096, keep_image=False)
    if not which or "ms" in which:
        fixture_multiscale(m)
    if not which or "nms" in which:
        fixture_nms()
    if not which or "match" in which:
        fixture_match()


if __name__ == "__main__":
    main(sys.argv[1:])
