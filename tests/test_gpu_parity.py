"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle and the
reference-generated golden fixtures.  Bit-exact for NMS / selection / match indices;
1e-3 (BASELINE.json north_star) for scores and descriptors."""
import os

import numpy as np
import pytest
import torch

from oracle import sfd2_oracle as orc
from sfd2_b200.synth import synth_image, synth_image_u8, synth_descriptors

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

TOL = 1e-3   # north_star: descriptors / scores within 1e-3


def _img(g):
    H, W = int(g["H"]), int(g["W"])
    if "image_u8" in g.files:
        return (g["image_u8"].astype(np.float32) / np.float32(255)).transpose(2, 0, 1)[None].copy()
    return synth_image(int(g["seed"]), H, W)


# ------------------------------------------------------------------ single conv layers
def _conv_ref(x, w, b, stride, groups, relu):
    import torch.nn.functional as F
    y = F.conv2d(torch.from_numpy(x.transpose(2, 0, 1)[None].copy()).double(), torch.from_numpy(w).double(),
                 torch.from_numpy(b).double(), stride=stride, padding=w.shape[2] // 2, groups=groups)
    return (F.relu(y) if relu else y)[0].permute(1, 2, 0).numpy()


CONV_CASES = [  # H, W, cin, cout, k, stride, groups, relu
    (40, 56, 64, 64, 3, 1, 1, 1), (41, 57, 64, 128, 3, 2, 1, 1), (24, 40, 128, 256, 3, 1, 1, 1),
    (24, 40, 256, 256, 1, 1, 1, 1), (40, 56, 256, 256, 3, 2, 1, 1), (30, 34, 256, 65, 3, 1, 1, 0),
    (30, 34, 256, 128, 3, 1, 1, 0), (24, 40, 256, 256, 3, 1, 32, 1), (17, 19, 64, 64, 3, 1, 1, 1),
    (41, 57, 64, 64, 3, 2, 1, 1),      # conv1b's shape class: stride 2, N = 64 ([w_hi | w_lo] single-MMA path in exact mode)
]


@pytest.mark.parametrize("prec,rtol", [("fp32", 2e-6), ("exact", 6e-6), ("fast", 3e-3)])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_layer(case, prec, rtol):
    from gpu_util import debug_conv
    H, W, cin, cout, k, stride, groups, relu = case
    rng = np.random.RandomState(1)
    x = (np.maximum(rng.randn(H, W, cin), 0) * 3).astype(np.float32)
    w = (rng.randn(cout, cin // groups, k, k) / np.sqrt(cin // groups * k * k)).astype(np.float32)
    b = (rng.randn(cout) * 0.1).astype(np.float32)
    y = debug_conv(x, w, b, stride, groups, relu, prec)
    ref = _conv_ref(x, w, b, stride, groups, relu)
    assert np.abs(y - ref).max() <= rtol * np.abs(ref).max()


# ------------------------------------------------------------------ NMS + selection (bit-exact)
def test_nms_cases_bit_exact(golden):
    from gpu_util import nms_select
    g = golden("nms_cases")
    for k in [f[3:] for f in g.files if f.startswith("in_")]:
        heat = g["in_" + k]
        xy, sc, out = nms_select(heat, conf_th=0.001, border=4, topk=8192)
        assert np.array_equal(out, g["out_" + k]), k
        rx, ry, rs = orc.select_keypoints(torch.from_numpy(g["out_" + k]), 0.001, 4, 8192)
        assert np.array_equal(xy[:, 0], rx) and np.array_equal(xy[:, 1], ry), k
        assert np.array_equal(sc, rs), k


@pytest.mark.parametrize("name", ["small_96x128", "odd_100x141"])
def test_nms_on_reference_heatmaps(golden, name):
    from gpu_util import nms_select
    g = golden(name)
    xy, sc, out = nms_select(g["heat"], conf_th=0.001, border=4, topk=int(g["K"]))
    assert np.array_equal(out, g["nms"])
    assert np.array_equal(xy.astype(np.int16), g["kp_xy"])
    assert np.array_equal(sc, g["scores"])


def test_topk_truncation_and_ties():
    from gpu_util import nms_select
    rng = np.random.RandomState(3)
    heat = np.zeros((200, 300), np.float32)
    ys, xs = np.meshgrid(np.arange(10, 190, 10), np.arange(10, 290, 10), indexing="ij")
    heat[ys, xs] = np.round(rng.rand(*ys.shape) * 4) / 8 + 0.125      # isolated peaks, many exact score ties
    for K in (1, 7, 100, 4096):
        xy, sc, _ = nms_select(heat, conf_th=0.001, border=4, topk=K)
        rx, ry, rs = orc.select_keypoints(orc.simple_nms(torch.from_numpy(heat)[None, None], 4), 0.001, 4, K)
        assert np.array_equal(xy[:, 0], rx) and np.array_equal(xy[:, 1], ry) and np.array_equal(sc, rs), K


def test_large_random_heatmap_full_size():
    """Full benchmark size: CUDA NMS+select against the oracle on a dense random map
    (every stage of the 3-round algorithm is exercised; ~30k survivors > smem sort capacity)."""
    from gpu_util import nms_select
    rng = np.random.RandomState(5)
    heat = (rng.rand(1200, 1600).astype(np.float32)) ** 2
    xy, sc, out = nms_select(heat, conf_th=0.001, border=4, topk=4096)
    ref = orc.simple_nms(torch.from_numpy(heat)[None, None], 4)
    assert np.array_equal(out, ref[0, 0].numpy())
    rx, ry, rs = orc.select_keypoints(ref, 0.001, 4, 4096)
    assert np.array_equal(xy[:, 0], rx) and np.array_equal(xy[:, 1], ry) and np.array_equal(sc, rs)


# ------------------------------------------------------------------ end-to-end extraction
def _log_count(key, value):
    """Measured parity counts go to gpurun_out/parity_counts.json (merged back from the GPU box) and to stdout."""
    import json
    print(f"[parity] {key} = {value}")
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, "parity_counts.json")
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[key] = value
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def _check_extract(out, g, exact_keypoints, score_rtol):
    """Keypoint parity.  `exact_keypoints`: the keypoint SET must equal the reference's, except
    candidates whose reference score is within `score_rtol` (the mode's measured arithmetic error)
    of the K-th / (K+1)-th score - there the cut itself is decided by noise smaller than the
    reference's own cuDNN-vs-oneDNN differences.  Order is checked up to swaps of near-equal scores."""
    kp = out["keypoints"].astype(np.int64)
    ref = g["kp_xy"].astype(np.int64)
    assert out["keypoints"].dtype == np.float64 and out["descriptors"].dtype == np.float64
    assert out["descriptors"].shape == (len(kp), 128) and out["scores"].shape == (len(kp),)
    assert len(kp) == len(ref)
    idx = {tuple(k): i for i, k in enumerate(ref)}
    hit = [(i, idx[tuple(k)]) for i, k in enumerate(kp) if tuple(k) in idx]
    i0, i1 = np.array(hit).T
    missing = len(ref) - len(hit)
    if exact_keypoints:
        cut = float(g["scores"][-1])
        nxt = float(g["next_score"])
        band = max(cut * score_rtol, 0.0)
        # every reference keypoint we miss (and every extra one we report) must sit in the tie band at the cut
        ours_only = [i for i, k in enumerate(kp) if tuple(k) not in idx]
        ref_only = sorted(set(range(len(ref))) - set(i1.tolist()))
        assert missing <= 2, f"{missing} keypoints differ"
        for j in ref_only:
            assert abs(float(g["scores"][j]) - cut) <= 4 * band + abs(cut - nxt), (j, g["scores"][j], cut)
        for i in ours_only:
            assert abs(out["scores"][i] - cut) <= 4 * band + abs(cut - nxt), (i, out["scores"][i], cut)
        # order: position may differ only among scores closer than the arithmetic error
        disp = np.nonzero(i0 != i1)[0]
        for d in disp:
            assert abs(g["scores"][i1[d]] - g["scores"][i0[d]]) <= 4 * score_rtol * g["scores"][i1[d]] + 1e-7, d
        assert np.abs(out["scores"][i0] - g["scores"][i1]).max() <= TOL
        assert np.abs(out["descriptors"][i0] - g["desc"][i1]).max() <= TOL
    else:
        # single-pass fp16: inside the 1e-3 tolerance but NOT keypoint-exact; a handful of pixels flip their
        # 3-class stability argmax (a discontinuity), which rescales that pixel's score by 2-10x
        assert len(hit) >= 0.98 * len(ref)
        ds = np.abs(out["scores"][i0] - g["scores"][i1])
        assert np.mean(ds <= TOL) >= 0.995 and np.median(ds) <= 1e-4
        assert np.abs(out["descriptors"][i0] - g["desc"][i1]).max() <= 2 * TOL
    assert np.all(np.diff(out["scores"]) <= 0)
    np.testing.assert_allclose(np.linalg.norm(out["descriptors"], axis=1), 1.0, atol=1e-5)
    return missing


@pytest.mark.parametrize("name", ["small_96x128", "odd_100x141", "c1_640x480", "c2_1600x1200"])
@pytest.mark.parametrize("prec", ["fp32", "exact", "mixed"])
def test_extract_matches_reference(golden, name, prec):
    """`mixed` = everything that feeds the heat-map in the 3-pass split, the descriptor head single-pass: the
    keypoint criteria are those of `exact`, descriptors must (only) meet the north-star 1e-3."""
    from gpu_util import model
    from sfd2_b200 import extract_resnet_return
    g = golden(name)
    out = extract_resnet_return(model(prec), torch.from_numpy(_img(g)), topK=int(g["K"]), conf_th=0.001, scales=[1.0])
    missing = _check_extract(out, g, exact_keypoints=True, score_rtol={"fp32": 1e-5, "exact": 1e-4, "mixed": 1e-4}[prec])
    _log_count(f"extract_missing[{prec}-{name}]", missing)
    # identical keypoint SETS on every fixture, C2 included (measured on B200: 0 missing in all three modes; the
    # relative gap at C2's cut is 2.2e-4 against an arithmetic error of ~1e-5)
    assert missing == 0, f"{missing} reference keypoints missing"


@pytest.mark.parametrize("name", ["c1_640x480", "c2_1600x1200"])
def test_extract_fast_mode_within_tolerance(golden, name):
    from gpu_util import model
    from sfd2_b200 import extract_resnet_return
    g = golden(name)
    out = extract_resnet_return(model("fast"), torch.from_numpy(_img(g)), topK=int(g["K"]), conf_th=0.001, scales=[1.0])
    _check_extract(out, g, exact_keypoints=False, score_rtol=3e-3)


def test_mixed_mode_keypoints_equal_exact_mode(golden):
    """The heat-map path of `mixed` is the `exact` path: keypoints and scores are bit-identical, only the
    descriptors differ (single-pass head), by less than the 1e-3 tolerance."""
    from gpu_util import model
    from sfd2_b200 import extract_resnet_return
    for name in ("c1_640x480", "c2_1600x1200"):
        g = golden(name)
        img = torch.from_numpy(_img(g))
        a = extract_resnet_return(model("exact"), img, topK=int(g["K"]), conf_th=0.001, scales=[1.0])
        b = extract_resnet_return(model("mixed"), img, topK=int(g["K"]), conf_th=0.001, scales=[1.0])
        assert np.array_equal(a["keypoints"], b["keypoints"]) and np.array_equal(a["scores"], b["scores"])
        d = np.abs(a["descriptors"] - b["descriptors"]).max()
        assert 0 < d <= TOL, d


def test_extract_device_and_u8_inputs_agree(golden):
    from gpu_util import model, WEIGHTS
    from sfd2_b200 import extract_resnet_return, Extractor
    g = golden("c1_640x480")
    img = torch.from_numpy(_img(g))
    a = extract_resnet_return(model("exact"), img, topK=1000, conf_th=0.001, scales=[1.0])
    b = extract_resnet_return(model("exact"), img.cuda(), topK=1000, conf_th=0.001, scales=[1.0])
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    ex = Extractor(WEIGHTS, use_stability=True, precision="exact", topk=1000)
    u8 = torch.from_numpy(np.stack([g["image_u8"], g["image_u8"]])).cuda()
    o = ex(u8)
    torch.cuda.synchronize()
    assert o["counts"].tolist() == [1000, 1000]
    assert np.array_equal(o["keypoints"][0].cpu().numpy(), a["keypoints"].astype(np.float32))
    assert np.array_equal(o["keypoints"][1].cpu().numpy(), a["keypoints"].astype(np.float32))
    assert np.abs(o["descriptors"][1].cpu().numpy() - a["descriptors"]).max() < 1e-6


def test_extract_edge_cases():
    from gpu_util import model
    from sfd2_b200 import extract_resnet_return
    m = model("exact")
    z = extract_resnet_return(m, torch.zeros(1, 3, 64, 80), topK=100, conf_th=0.5, scales=[1.0])   # nothing above 0.5
    assert z["keypoints"].shape == (0, 2) and z["descriptors"].shape == (0, 128) and z["scores"].shape == (0,)
    img = torch.from_numpy(synth_image(9, 72, 88))
    full = extract_resnet_return(m, img, topK=-1, conf_th=0.001, scales=[1.0])
    one = extract_resnet_return(m, img, topK=1, conf_th=0.001, scales=[1.0])
    assert one["keypoints"].shape == (1, 2) and np.array_equal(one["keypoints"][0], full["keypoints"][0])
    ref = orc.extract(orc.load_state(os.path.join(os.path.dirname(__file__), "..", "weights", "ressegnetv2_wapv2.npz")),
                      img.numpy(), topK=-1)
    assert np.array_equal(full["keypoints"], ref["keypoints"])
    assert np.abs(full["descriptors"] - ref["descriptors"]).max() <= TOL


# ------------------------------------------------------------------ matchers
@pytest.mark.parametrize("prec", ["fp32", "exact"])
def test_matchers_match_reference(golden, prec):
    from sfd2_b200 import NearestNeighbor, Matcher, matcher_confs
    g = golden("match_cases")
    for tag in ["sq", "wide", "tall", "one", "col"]:
        d0, d1 = g[f"{tag}_d0"], g[f"{tag}_d1"]
        data = {"descriptors0": torch.from_numpy(d0.T.copy())[None].cuda(),
                "descriptors1": torch.from_numpy(d1.T.copy())[None].cuda()}
        out = NearestNeighbor({"do_mutual_check": True, "precision": prec})(data)
        assert out["matches0"].dtype == torch.int64 and out["matches0"].shape == (1, len(d0))
        assert np.array_equal(out["matches0"][0].cpu().numpy(), g[f"{tag}_hloc_m0"]), tag
        assert np.abs(out["matching_scores0"][0].cpu().numpy() - g[f"{tag}_hloc_s0"]).max() <= TOL
        o1 = NearestNeighbor({"do_mutual_check": False, "precision": prec})(data)
        assert np.array_equal(o1["matches0"][0].cpu().numpy(), g[f"{tag}_hloc_nomutual_m0"]), tag
        if f"{tag}_itloc_m0" in g.files:
            o2 = Matcher(matcher_confs["NNM"], precision=prec)({"descriptors0": d0.astype(np.float64),
                                                               "descriptors1": d1.astype(np.float64)})
            assert np.array_equal(o2["matches0"], g[f"{tag}_itloc_m0"]), tag
            assert np.abs(o2["matching_scores0"] - g[f"{tag}_itloc_s0"]).max() <= TOL
            o2["matches0"][0] = 5     # callers mutate the result in place (localize_cv2.py:557-559)


@pytest.mark.parametrize("name", ["c1_640x480", "c2_1600x1200"])
@pytest.mark.parametrize("prec", ["fp32", "exact"])
def test_pair_matches_reference(golden, name, prec):
    from sfd2_b200 import NearestNeighbor
    g = golden(name)
    data = {"descriptors0": torch.from_numpy(g["desc"].T.copy())[None].cuda(),
            "descriptors1": torch.from_numpy(g["desc_b"].T.copy())[None].cuda()}
    out = NearestNeighbor({"do_mutual_check": True, "precision": prec})(data)
    m0 = out["matches0"][0].cpu().numpy()
    ref = g["hloc_matches0"]
    if not np.array_equal(m0, ref):
        # disagreements are only allowed where the fp64 top-1/top-2 gap is below fp32 resolution
        nn12, nn21, gap, _ = orc.mutual_nn_exact(g["desc"], g["desc_b"])
        _, _, cgap, _ = orc.mutual_nn_exact(g["desc_b"], g["desc"])
        bad = np.nonzero(m0 != ref)[0]
        assert len(bad) <= 2, f"{len(bad)} rows differ"
        for i in bad:
            # the row's own arg-max is near-tied, or the mutual check of a column it involves is
            cols = {int(c) for c in (m0[i], ref[i], nn12[i]) if c >= 0}
            assert gap[i] < 1e-6 or any(cgap[c] < 1e-6 for c in cols), (int(i), float(gap[i]), [float(cgap[c]) for c in cols])
    assert np.abs(out["matching_scores0"][0].cpu().numpy() - g["hloc_scores0"]).max() <= TOL


def test_matcher_edge_cases():
    from sfd2_b200.matchers import match_dev, match_batched
    d0, d1 = synth_descriptors(4, 257, 130)
    a, b = torch.from_numpy(d0).cuda(), torch.from_numpy(d1).cuda()
    m0, s0 = match_dev(a, b[:0], precision="exact")          # empty db
    assert (m0 == -1).all()
    m0, s0 = match_dev(a[:0], b, precision="exact")          # empty query
    assert m0.numel() == 0
    dup = torch.cat([b, b])                                   # duplicated columns: lowest index wins
    m1, _ = match_dev(a, dup, mutual=False, precision="exact")
    assert int(m1.max()) < len(d1)
    # full size, size-independent property: mutual matches are symmetric
    e0, e1 = synth_descriptors(6, 4096, 4096)
    x, y = torch.from_numpy(e0).cuda(), torch.from_numpy(e1).cuda()
    f, _ = match_dev(x, y, precision="exact")
    r, _ = match_dev(y, x, precision="exact")
    f, r = f.cpu().numpy(), r.cpu().numpy()
    ok = f >= 0
    assert ok.sum() > 1500 and np.array_equal(r[f[ok]], np.nonzero(ok)[0])
    ref = orc.match_hloc(e0.T[None], e1.T[None])["matches0"][0].numpy()
    assert (f == ref).mean() == 1.0
    # batched API == per-pair API
    off0, off1 = [0, 100, 257], [0, 60, 130]
    mb, sb = match_batched(a, off0, b, off1, precision="exact")
    for i in range(2):
        mi, si = match_dev(a[off0[i]:off0[i + 1]], b[off1[i]:off1[i + 1]], precision="exact")
        assert torch.equal(mb[off0[i]:off0[i + 1]], mi) and torch.equal(sb[off0[i]:off0[i + 1]], si)


@pytest.mark.parametrize("prec", ["fp32", "exact"])
def test_ratio_tests_match_reference(golden, prec):
    """Lowe ratio + distance thresholds (hloc find_nn) and it_loc 'nnr' against reference fixtures.
    Rows whose ratio sits within 1e-5 of the threshold may legitimately flip (fp32 GEMM order)."""
    from sfd2_b200 import NearestNeighbor, Matcher, matcher_confs
    g = golden("match_cases")
    c1 = golden("c1_640x480")
    cases = {t: (g[f"{t}_d0"], g[f"{t}_d1"]) for t in ["sq", "wide", "tall"]}
    cases["c1"] = (c1["desc"], c1["desc_b"])
    for tag, (d0, d1) in cases.items():
        data = {"descriptors0": torch.from_numpy(d0.T.copy())[None].cuda(),
                "descriptors1": torch.from_numpy(d1.T.copy())[None].cuda()}
        o = NearestNeighbor({"do_mutual_check": True, "ratio_threshold": 0.8, "distance_threshold": 0.7,
                             "precision": prec})(data)
        m = o["matches0"][0].cpu().numpy()
        assert (m != g[f"{tag}_hloc_ratio_m0"]).sum() <= 1, (tag, (m != g[f"{tag}_hloc_ratio_m0"]).sum())
        ds = np.abs(o["matching_scores0"][0].cpu().numpy() - g[f"{tag}_hloc_ratio_s0"])
        assert (ds > TOL).sum() <= 1, tag
        o1 = NearestNeighbor({"do_mutual_check": False, "ratio_threshold": 0.9, "precision": prec})(data)
        m1 = o1["matches0"][0].cpu().numpy()
        assert (m1 != g[f"{tag}_hloc_ratio_nomutual_m0"]).sum() <= 1, tag
        o2 = Matcher(matcher_confs["NNR"], precision=prec)({"descriptors0": d0.astype(np.float64),
                                                            "descriptors1": d1.astype(np.float64)})
        assert (o2["matches0"] != g[f"{tag}_itloc_nnr_m0"]).sum() <= 1, tag


@pytest.mark.parametrize("prec", ["fp32", "exact"])
def test_multiscale_extract_matches_reference(golden, prec):
    """scales=[1.0, 0.75, 1.25] (nets/extractor.py:113-125): per-scale bilinear resize, border test against the
    original extents, keypoints mapped back, union cut to topK."""
    from gpu_util import model
    from sfd2_b200 import extract_resnet_return
    g = golden("ms_128x160")
    img = torch.from_numpy((g["image_u8"].astype(np.float32) / np.float32(255)).transpose(2, 0, 1)[None].copy())
    out = extract_resnet_return(model(prec), img, topK=int(g["K"]), conf_th=0.001, scales=list(g["scales"]))
    # the same (x, y) can be reported by several scales with different scores: match on (x, y, ~score)
    ref_kp, ref_sc = g["kp"].astype(np.float64), g["scores"].astype(np.float64)
    used, hit = set(), []
    for i, (k, s_) in enumerate(zip(out["keypoints"], out["scores"])):
        cand = [j for j in np.nonzero(np.all(np.abs(ref_kp - k) < 1e-3, axis=1))[0] if j not in used]
        if cand:
            j = min(cand, key=lambda j: abs(ref_sc[j] - s_))
            used.add(j)
            hit.append((i, j))
    assert len(hit) >= len(ref_sc) - 2, f"{len(hit)}/{len(ref_sc)} keypoints shared"
    i0, i1 = np.array(hit).T
    assert np.abs(out["scores"][i0] - g["scores"][i1]).max() <= TOL
    assert np.abs(out["descriptors"][i0] - g["desc"][i1]).max() <= TOL


def test_real_world_size_against_oracle(oracle_state):
    """1600x1063 (an Aachen image resized to max side 1600): H is odd and not a multiple of 8, so every stride-2
    layer uses ceil sizes, the heat-map comes out 1064 rows and is bilinearly resized to 1063 before NMS
    (nets/extractor.py:137-138).  CUDA (exact mode) against the CPU oracle."""
    from gpu_util import model
    from sfd2_b200 import extract_resnet_return
    img = synth_image(21, 1063, 1600)
    ref = orc.extract(oracle_state, img, topK=2048, conf_th=0.001)
    out = extract_resnet_return(model("exact"), torch.from_numpy(img), topK=2048, conf_th=0.001, scales=[1.0])
    a = set(map(tuple, out["keypoints"].astype(int)))
    b = set(map(tuple, ref["keypoints"].astype(int)))
    assert len(a & b) >= len(b) - 2, f"{len(a & b)}/{len(b)} keypoints shared"
    idx = {tuple(k): i for i, k in enumerate(ref["keypoints"].astype(int))}
    hit = [(i, idx[tuple(k)]) for i, k in enumerate(out["keypoints"].astype(int)) if tuple(k) in idx]
    i0, i1 = np.array(hit).T
    assert np.abs(out["scores"][i0] - ref["scores"][i1]).max() <= TOL
    assert np.abs(out["descriptors"][i0] - ref["descriptors"][i1]).max() <= TOL


@pytest.mark.parametrize("prec", ["fp32", "exact"])
def test_one_to_many_matches_pairwise(prec):
    """The grouped one-query-vs-many-db launch must equal the per-pair calls (ragged db sizes, incl. tiny ones)."""
    from sfd2_b200.matchers import match_dev, match_one_to_many
    rng = np.random.RandomState(8)
    q, _ = synth_descriptors(31, 1000, 10)
    sizes = [700, 128, 1, 333, 2049, 64]
    dbs = []
    for k, m in enumerate(sizes):
        d = rng.randn(m, 128).astype(np.float32)
        take = min(m, 300)
        d[:take] = q[rng.permutation(1000)[:take]] + 0.2 * rng.randn(take, 128).astype(np.float32)
        dbs.append(d / np.linalg.norm(d, axis=1, keepdims=True))
    off = np.concatenate([[0], np.cumsum(sizes)])
    qd = torch.from_numpy(q).cuda()
    dbd = torch.from_numpy(np.concatenate(dbs)).cuda()
    m_all, s_all = match_one_to_many(qd, dbd, off, precision=prec)
    assert m_all.shape == (len(sizes), 1000)
    for k in range(len(sizes)):
        m, s_ = match_dev(qd, dbd[off[k]:off[k + 1]], precision=prec)
        assert torch.equal(m_all[k], m), k
        assert torch.allclose(s_all[k], s_, atol=1e-6), k
        ref = orc.match_hloc(q.T[None], dbs[k].T[None])["matches0"][0].numpy()
        assert (m.cpu().numpy() == ref).mean() > 0.999


@pytest.mark.parametrize("hw", [(16, 16), (17, 33), (40, 24), (64, 64), (200, 136)])
def test_small_and_ragged_sizes_against_oracle(oracle_state, hw):
    """Tiny / ragged images: single-tile layers, clusters with one real tile, TMA boxes larger than the map."""
    from gpu_util import model
    from sfd2_b200 import extract_resnet_return
    H, W = hw
    img = synth_image(40 + H, H, W, sigma=1.5)
    ref = orc.extract(oracle_state, img, topK=500, conf_th=0.0005)
    for prec in ("fp32", "exact"):
        out = extract_resnet_return(model(prec), torch.from_numpy(img), topK=500, conf_th=0.0005, scales=[1.0])
        a = set(map(tuple, out["keypoints"].astype(int)))
        b = set(map(tuple, ref["keypoints"].astype(int)))
        assert len(a ^ b) <= 1, (prec, hw, len(a), len(b))
        # the dense heat-map too (tiny images yield few or no keypoints); isolated stability-class flips excepted
        heat = model(prec).debug_fetch("heat", (H, W))
        hm, _ = orc.heatmap(oracle_state, torch.from_numpy(img))
        err = np.abs(heat - hm[0, 0].numpy())
        assert np.mean(err <= 5e-5) >= 0.999 and np.median(err) <= 1e-5, (prec, hw, float(err.max()))
        if len(b):
            idx = {tuple(k): i for i, k in enumerate(ref["keypoints"].astype(int))}
            hit = [(i, idx[tuple(k)]) for i, k in enumerate(out["keypoints"].astype(int)) if tuple(k) in idx]
            i0, i1 = np.array(hit).T
            assert np.abs(out["scores"][i0] - ref["scores"][i1]).max() <= TOL
            assert np.abs(out["descriptors"][i0] - ref["descriptors"][i1]).max() <= TOL


def test_size_switching_and_context_lifecycle():
    """The workspace is re-created when the image size changes; results must not depend on call history,
    and contexts can be created / destroyed repeatedly."""
    from gpu_util import WEIGHTS
    from sfd2_b200 import get_model, extract_resnet_return
    a_img = torch.from_numpy(synth_image(50, 120, 160))
    b_img = torch.from_numpy(synth_image(51, 96, 200))
    first = None
    for rep in range(3):
        m, _ = get_model("ressegnetv2", WEIGHTS, use_stability=True)
        m.cuda()
        ra = extract_resnet_return(m, a_img, topK=200, conf_th=0.001, scales=[1.0])
        rb = extract_resnet_return(m, b_img, topK=200, conf_th=0.001, scales=[1.0])
        ra2 = extract_resnet_return(m, a_img, topK=200, conf_th=0.001, scales=[1.0])
        for k in ra:
            assert np.array_equal(ra[k], ra2[k]), k
        if first is None:
            first = (ra, rb)
        else:
            for k in ra:
                assert np.array_equal(ra[k], first[0][k]) and np.array_equal(rb[k], first[1][k]), k
        m.ctx.close()


def test_large_image_modes_agree():
    """2048 x 2560 (larger than the benchmark size): tcgen05 exact mode against the CUDA-core fp32 mode."""
    from gpu_util import model
    from sfd2_b200 import extract_resnet_return
    img = torch.from_numpy(synth_image(60, 2048, 2560))
    a = extract_resnet_return(model("fp32"), img, topK=8192, conf_th=0.001, scales=[1.0])
    b = extract_resnet_return(model("exact"), img, topK=8192, conf_th=0.001, scales=[1.0])
    ka, kb = set(map(tuple, a["keypoints"].astype(int))), set(map(tuple, b["keypoints"].astype(int)))
    assert len(ka) == 8192 and len(ka & kb) >= 8192 - 8
    idx = {tuple(k): i for i, k in enumerate(a["keypoints"].astype(int))}
    hit = [(i, idx[tuple(k)]) for i, k in enumerate(b["keypoints"].astype(int)) if tuple(k) in idx]
    i0, i1 = np.array(hit).T
    assert np.abs(b["scores"][i0] - a["scores"][i1]).max() <= TOL
    assert np.abs(b["descriptors"][i0] - a["descriptors"][i1]).max() <= TOL


# ------------------------------------------------------------------ the batched paths the bench numbers come from
def _c2_batch(golden, n=8):
    g = golden("c2_1600x1200")
    H, W = int(g["H"]), int(g["W"])
    u8 = [synth_image_u8(int(g["seed"]) + i, H, W) for i in range(n)]
    f32 = np.stack([(u.astype(np.float32) / np.float32(255)).transpose(2, 0, 1) for u in u8])
    return g, u8, f32


@pytest.mark.parametrize("prec", ["mixed", "exact"])
def test_batched_paths_equal_single_image_calls(golden, prec):
    """8 DISTINCT 1600x1200 images through (a) Extractor.extract_host (sfd2_extract_host, n = 8: copy stream, one event
    per image, workspaces alternating on two internal streams) and (b) Extractor.__call__ (sfd2_extract_dev, n = 8)
    must equal, image by image and bit for bit, the reference-signature call extract_resnet_return on that image
    alone - and image 0 must equal the reference fixture.  Run twice so workspace reuse across calls is covered."""
    from gpu_util import model, WEIGHTS
    from sfd2_b200 import extract_resnet_return, Extractor
    g, u8, f32 = _c2_batch(golden)
    K = int(g["K"])
    single = [extract_resnet_return(model(prec), torch.from_numpy(f32[i:i + 1]), topK=K, conf_th=0.001, scales=[1.0])
              for i in range(len(f32))]
    missing = _check_extract(single[0], g, exact_keypoints=True, score_rtol=1e-4)
    assert missing == 0
    ex = Extractor(WEIGHTS, use_stability=True, precision=prec, topk=K, conf_th=0.001)
    host = torch.from_numpy(f32).pin_memory()
    dev = torch.from_numpy(f32).cuda()
    dev_u8 = torch.from_numpy(np.stack(u8)).cuda()
    for rep in range(2):
        h = ex.extract_host(host)
        d = ex(dev)
        ex.check_status()
        d8 = ex(dev_u8)
        ex.check_status()
        for i, ref in enumerate(single):
            n = len(ref["scores"])
            for name, o in (("host", h), ("dev", {k: v.cpu().numpy() for k, v in d.items()}),
                            ("dev_u8", {k: v.cpu().numpy() for k, v in d8.items()})):
                assert int(o["counts"][i]) == n, (name, rep, i)
                assert np.array_equal(o["keypoints"][i, :n].astype(np.float64), ref["keypoints"]), (name, rep, i)
                assert np.array_equal(o["scores"][i, :n].astype(np.float64), ref["scores"]), (name, rep, i)
                assert np.array_equal(o["descriptors"][i, :n].astype(np.float64), ref["descriptors"]), (name, rep, i)
    # distinct images really gave distinct results
    assert not np.array_equal(single[0]["keypoints"], single[1]["keypoints"])


def test_pair_pipeline_matches_reference(golden):
    """BASELINE configs[4] (C5): extract G(s), extract roll(G(s)), mutual-NN match - through the sweep the bench times
    (sfd2_b200.sweep.pair_sweep: batched device extraction + one match per pair) against the reference fixture of the
    same pair (hloc/match_features.py:90-121 on the reference's own features)."""
    from gpu_util import WEIGHTS
    from sfd2_b200.sweep import pair_sweep
    from sfd2_b200.synth import shifted_twin
    g = golden("c2_1600x1200")
    H, W, K = int(g["H"]), int(g["W"]), int(g["K"])
    a = synth_image_u8(int(g["seed"]), H, W)
    pairs = [(a, shifted_twin(a)), (synth_image_u8(77, H, W), shifted_twin(synth_image_u8(77, H, W)))]
    res = pair_sweep(pairs, WEIGHTS, precision="mixed", topk=K, conf_th=0.001, keep=True)
    r0 = res["pairs"][0]
    # same keypoint SETS as the reference on both frames (order may differ among near-equal scores): translate our row /
    # column numbering into the fixture's before comparing matches
    def to_ref(ours, ref):
        idx = {tuple(k): i for i, k in enumerate(ref.astype(np.int64))}
        perm = np.array([idx.get(tuple(k), -1) for k in ours.astype(np.int64)])
        assert (perm >= 0).all() and len(set(perm.tolist())) == len(ref), "keypoint sets differ from the reference"
        return perm
    p0, p1 = to_ref(r0["keypoints0"], g["kp_xy"]), to_ref(r0["keypoints1"], g["kp_xy_b"])
    _log_count("pair_pipeline_rows_out_of_order[mixed-c2]", int((p0 != np.arange(len(p0))).sum()))
    m0 = np.full(len(p0), -1, np.int64)
    ours = r0["matches0"]
    m0[p0] = np.where(ours >= 0, p1[np.maximum(ours, 0)], -1)
    sim = np.zeros(len(p0), np.float32)
    sim[p0] = r0["sim0"]
    bad = np.nonzero(m0 != g["hloc_matches0"])[0]
    _log_count("pair_pipeline_rows_differing[mixed-c2]", int(len(bad)))
    # our descriptors differ from the reference's by <= 1e-3 (single-pass head), so a near-tied arg-max may flip
    nn12, nn21, gap, _ = orc.mutual_nn_exact(g["desc"], g["desc_b"])
    _, _, cgap, _ = orc.mutual_nn_exact(g["desc_b"], g["desc"])
    for i in bad:
        cols = {int(c) for c in (m0[i], g["hloc_matches0"][i], nn12[i]) if c >= 0}
        assert gap[i] < 4e-3 or any(cgap[c] < 4e-3 for c in cols), (int(i), float(gap[i]))
    assert len(bad) <= 0.01 * len(m0)
    assert np.abs((sim + 1) / 2 - g["hloc_scores0"]).max() <= TOL
    assert res["pairs"][1]["matches0"].shape == (len(res["pairs"][1]["keypoints0"]),)
    assert (res["pairs"][1]["matches0"] >= 0).sum() > 1000


# ------------------------------------------------------------------ grouped matcher: device-side counts, layouts, ids
def test_grouped_pairs_with_device_counts_equal_single_calls():
    """sfd2_match_pairs_dev on fixed-capacity sets whose valid row counts live on the device (what the extract -> match
    pipeline feeds it) must equal per-pair calls on the trimmed sets, for ragged counts incl. 0, 1 and the capacity."""
    from sfd2_b200.matchers import match_dev, match_pairs_dev
    K = 700
    counts = [700, 513, 128, 1, 0, 257, 640, 699]
    rng = np.random.RandomState(11)
    base, _ = synth_descriptors(12, 900, 10)
    D = np.zeros((len(counts), K, 128), np.float32)
    for i, c in enumerate(counts):
        d = base[rng.permutation(900)[:K]] + 0.25 * rng.randn(K, 128).astype(np.float32)
        D[i] = d / np.linalg.norm(d, axis=1, keepdims=True)
    dd = torch.from_numpy(D).cuda()
    cc = torch.tensor(counts, dtype=torch.int32, device="cuda")
    idx0 = [0, 1, 2, 3, 4, 5, 6, 0, 7]
    idx1 = [1, 2, 0, 0, 1, 6, 5, 4, 7]
    for mutual in (True, False):
        m, s_ = match_pairs_dev(dd, cc, idx0, idx1, mutual=mutual, precision="exact")
        assert m.shape == (len(idx0), K)
        for p, (a, b) in enumerate(zip(idx0, idx1)):
            na, nb = counts[a], counts[b]
            mi, si = match_dev(dd[a, :na], dd[b, :nb], mutual=mutual, precision="exact")
            assert torch.equal(m[p, :na], mi), (mutual, p)
            assert torch.equal(s_[p, :na], si), (mutual, p)
            assert bool((m[p, na:] == -1).all()) and bool((s_[p, na:] == 0).all()), (mutual, p)
            if na and nb:
                ref = orc.match_hloc(D[a, :na].T[None], D[b, :nb].T[None], do_mutual_check=mutual)["matches0"][0].numpy()
                assert (mi.cpu().numpy() == ref).mean() > 0.999, (mutual, p)


def test_hloc_layout_is_read_in_place(golden):
    """[1, D, N] (nearest_neighbor.py:39) goes to the kernels as is; results equal the row-major call bit for bit."""
    from sfd2_b200.matchers import match_dev
    g = golden("match_cases")
    for tag in ["sq", "wide", "tall", "one", "col"]:
        d0, d1 = g[f"{tag}_d0"], g[f"{tag}_d1"]
        a, b = torch.from_numpy(d0).cuda(), torch.from_numpy(d1).cuda()
        at, bt = torch.from_numpy(d0.T.copy()).cuda(), torch.from_numpy(d1.T.copy()).cuda()
        for kw in ({}, {"ratio_th": 0.8, "dist_th": 0.7}, {"mutual": False}):
            m_r, s_r = match_dev(a, b, precision="exact", **kw)
            m_c, s_c = match_dev(at, bt, precision="exact", layout="cols", **kw)
            assert torch.equal(m_r, m_c) and torch.equal(s_r, s_c), (tag, kw)


def test_localizer_subset_and_remap_on_device():
    """it_loc/localize_cv2.py:511-560: match a query against db images using only db keypoints with a 3-D point
    (db_3D_ids != -1) and report ORIGINAL db rows - one grouped launch for the whole list of db images."""
    from sfd2_b200.matchers import feature_matching
    rng = np.random.RandomState(21)
    q, _ = synth_descriptors(41, 1500, 10)
    sizes = [900, 300, 2, 1201, 5, 128]
    dbs, ids = [], []
    for k, m in enumerate(sizes):
        d = rng.randn(m, 128).astype(np.float32)
        take = min(m, 400)
        d[:take] = q[rng.permutation(1500)[:take]] + 0.2 * rng.randn(take, 128).astype(np.float32)
        dbs.append(d / np.linalg.norm(d, axis=1, keepdims=True))
        i3 = rng.randint(0, 10000, m)
        i3[rng.rand(m) < (0.5 if k != 4 else 0.9)] = -1
        ids.append(i3)
    ids[2][:] = -1                                   # nothing valid
    ids[5][:] = np.arange(128)                       # everything valid
    out = feature_matching(q.astype(np.float64), dbs, db_3D_ids=ids)
    for k in range(len(sizes)):
        ref = orc.feature_matching(q, dbs[k], ids[k])
        assert out[k].shape == ref.shape
        assert (out[k] == ref).mean() > 0.999, (k, (out[k] != ref).sum())
        ok = out[k] >= 0
        assert (ids[k][out[k][ok]] != -1).all(), k     # only keypoints with a 3-D point are ever matched
    single = feature_matching(q, dbs[0], db_3D_ids=ids[0])
    assert np.array_equal(single, out[0])
    plain = feature_matching(q, dbs[0])
    assert (plain == orc.feature_matching(q, dbs[0])).mean() > 0.999


def test_prefetch_overlaps_upload_without_changing_results(golden):
    """model.prefetch(next_image) + extract_resnet_return(model, next_image) == the plain call, for the image that was
    prefetched; a call with a different tensor ignores the prefetched one."""
    from gpu_util import model
    from sfd2_b200 import extract_resnet_return
    m = model("mixed")
    imgs = [torch.from_numpy(synth_image(70 + i, 240, 320)).pin_memory() for i in range(3)]
    plain = [extract_resnet_return(m, im, topK=500, conf_th=0.001, scales=[1.0]) for im in imgs]
    m.prefetch(imgs[0])
    for i in range(3):
        if i + 1 < 3:
            cur = extract_resnet_return(m, imgs[i], topK=500, conf_th=0.001, scales=[1.0])    # consumes the prefetched copy
            m.prefetch(imgs[i + 1])
        else:
            cur = extract_resnet_return(m, imgs[i], topK=500, conf_th=0.001, scales=[1.0])
        for k in cur:
            assert np.array_equal(cur[k], plain[i][k]), (i, k)
    m.prefetch(imgs[0])
    other = extract_resnet_return(m, imgs[2], topK=500, conf_th=0.001, scales=[1.0])           # not the prefetched tensor
    for k in other:
        assert np.array_equal(other[k], plain[2][k]), k


def test_candidate_overflow_is_reported_on_every_path(monkeypatch):
    """More NMS candidates than the workspace holds (H*W/16 + 4096; cannot happen without exact plateaus, so the test
    hook SFD2_CAND_CAP shrinks the list): the synchronous calls must raise - host path AND device path - the batched
    device path must report it through check_status(), and the flag must not leak into the next call.  The standalone
    NMS entry point is exercised with a real plateau."""
    from gpu_util import WEIGHTS, nms_select
    from sfd2_b200 import get_model, extract_resnet_return, Extractor, _lib
    with pytest.raises(_lib.Sfd2Error, match="overflow"):
        nms_select(np.full((200, 300), 0.5, np.float32), conf_th=0.001, border=4, topk=100)      # every pixel ties
    monkeypatch.setenv("SFD2_CAND_CAP", "64")
    m, _ = get_model("ressegnetv2", WEIGHTS, use_stability=True, precision="exact")
    m.cuda()
    ex = Extractor(WEIGHTS, precision="exact", topk=32, conf_th=0.001)
    monkeypatch.delenv("SFD2_CAND_CAP")
    img = torch.from_numpy(synth_image(9, 160, 200))                     # ~100 candidates > 64
    with pytest.raises(_lib.Sfd2Error, match="candidates"):
        extract_resnet_return(m, img, topK=32, conf_th=0.001, scales=[1.0])
    with pytest.raises(_lib.Sfd2Error, match="candidates"):
        extract_resnet_return(m, img.cuda(), topK=32, conf_th=0.001, scales=[1.0])
    few = extract_resnet_return(m, img, topK=32, conf_th=0.05, scales=[1.0])     # few candidates: fine, and the flag was cleared
    assert len(few["scores"]) <= 32
    ex(img.cuda())
    with pytest.raises(_lib.Sfd2Error, match="candidates"):
        ex.check_status()
    ex.check_status()                                                     # cleared by the failing query


def test_matcher_fuzz_against_fp64():
    """Random shapes (1 .. 1700 rows, tile-boundary and prime sizes), all modes, single calls and grouped calls with
    device-side counts: every row must equal the float64 evaluation unless its decision is near-tied (< 2e-6)."""
    from sfd2_b200.matchers import match_dev, match_sets_dev
    rng = np.random.RandomState(123)
    sizes = [1, 2, 127, 128, 129, 255, 256, 257, 383, 511, 640, 769, 1021, 1279, 1700]

    def make(n, seed):
        base, _ = synth_descriptors(seed % 7, 1800, 10)
        d = base[rng.permutation(1800)[:n]] + 0.3 * rng.randn(n, 128).astype(np.float32)
        return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)

    def expect(d0, d1, mutual, ratio):
        sim = d0.astype(np.float64) @ d1.astype(np.float64).T
        nn12 = sim.argmax(1)
        ok = np.ones(len(d0), bool)
        srt = np.sort(sim, axis=1)
        gap = srt[:, -1] - srt[:, -2] if sim.shape[1] > 1 else np.full(len(d0), np.inf)
        cs = np.sort(sim, axis=0)
        cgap = cs[-1] - cs[-2] if sim.shape[0] > 1 else np.full(sim.shape[1], np.inf)
        if ratio:
            if sim.shape[1] > 1:
                ok &= 2 * (1 - srt[:, -1]) <= ratio * ratio * 2 * (1 - srt[:, -2])
        if mutual:
            nn21 = sim.argmax(0)
            ok &= nn21[nn12] == np.arange(len(d0))
            if ratio and sim.shape[0] > 1:
                ok &= (2 * (1 - cs[-1]) <= ratio * ratio * 2 * (1 - cs[-2]))[nn12]
        return np.where(ok, nn12, -1), gap, cgap, nn12

    ncase = 0
    for it in range(24):
        n0, n1 = int(rng.choice(sizes)), int(rng.choice(sizes))
        d0, d1 = make(n0, it), make(n1, it + 100)
        a, b = torch.from_numpy(d0).cuda(), torch.from_numpy(d1).cuda()
        for mutual, ratio in ((True, None), (False, None), (True, 0.9)):
            m, _ = match_dev(a, b, mutual=mutual, ratio_th=ratio, precision="exact")
            ref, gap, cgap, nn12 = expect(d0, d1, mutual, ratio)
            bad = np.nonzero(m.cpu().numpy() != ref)[0]
            for i in bad:
                near = gap[i] < 2e-6 or cgap[nn12[i]] < 2e-6 or ratio is not None
                assert near, (n0, n1, mutual, ratio, int(i), float(gap[i]))
            assert len(bad) <= max(2, n0 // 100), (n0, n1, mutual, ratio, len(bad))
            ncase += 1
    # grouped call: capacity 1700 per set, valid counts on the device
    K = 1700
    counts = [int(rng.choice(sizes)) for _ in range(10)]
    D = np.zeros((len(counts), K, 128), np.float32)
    for i, c in enumerate(counts):
        D[i, :c] = make(c, 50 + i)
    dd, cc = torch.from_numpy(D).cuda(), torch.tensor(counts, dtype=torch.int32, device="cuda")
    sets = [{"data": dd[i], "count": cc[i:i + 1]} for i in range(len(counts))]
    pa = [int(x) for x in rng.randint(0, len(counts), 16)]
    pb = [int(x) for x in rng.randint(0, len(counts), 16)]
    m, _ = match_sets_dev(sets, pa, pb, precision="exact")
    m = m.view(16, K).cpu().numpy()
    for k, (ia, ib) in enumerate(zip(pa, pb)):
        ref, gap, cgap, nn12 = expect(D[ia, :counts[ia]], D[ib, :counts[ib]], True, None)
        got = m[k, :counts[ia]]
        bad = np.nonzero(got != ref)[0]
        for i in bad:
            assert gap[i] < 2e-6 or cgap[nn12[i]] < 2e-6, (k, ia, ib, int(i))
        assert (m[k, counts[ia]:] == -1).all()
    assert ncase == 72


@pytest.mark.parametrize("prec", ["mixed", "fast"])
def test_sparse_descriptor_head_equals_dense(golden, monkeypatch, prec):
    """Single-pass modes evaluate the descriptor head only at the sampled tap pixels (tc_desc_sparse.cu): every output must
    be bit-identical to the dense head + sample_kernel path (SFD2_SPARSE_DESC=0), incl. odd sizes, K > candidates, 0 keypoints
    and the batched path."""
    from gpu_util import WEIGHTS
    from sfd2_b200 import get_model, extract_resnet_return, Extractor
    monkeypatch.setenv("SFD2_SPARSE_DESC", "0")
    dense, _ = get_model("ressegnetv2", WEIGHTS, use_stability=True, precision=prec)
    dense.cuda()
    exd = Extractor(WEIGHTS, precision=prec, topk=700)
    monkeypatch.delenv("SFD2_SPARSE_DESC")
    sparse, _ = get_model("ressegnetv2", WEIGHTS, use_stability=True, precision=prec)
    sparse.cuda()
    exs = Extractor(WEIGHTS, precision=prec, topk=700)
    imgs = [_img(golden("c1_640x480")), _img(golden("odd_100x141")), synth_image(33, 250, 333), synth_image(34, 64, 80, sigma=1.5)]
    for im in imgs:
        for K in (300, 5000, -1):
            a = extract_resnet_return(dense, torch.from_numpy(im), topK=K, conf_th=0.001, scales=[1.0])
            b = extract_resnet_return(sparse, torch.from_numpy(im), topK=K, conf_th=0.001, scales=[1.0])
            for k in a:
                assert np.array_equal(a[k], b[k]), (im.shape, K, k, float(np.abs(a[k] - b[k]).max()) if a[k].shape == b[k].shape else None)
    z = extract_resnet_return(sparse, torch.zeros(1, 3, 64, 80), topK=100, conf_th=0.5, scales=[1.0])
    assert z["descriptors"].shape == (0, 128)
    batch = torch.from_numpy(np.concatenate([synth_image(40 + i, 240, 320) for i in range(5)])).cuda()
    od, os_ = exd(batch), exs(batch)
    for k in od:
        assert torch.equal(od[k], os_[k]), k


def test_single_host_image_band_upload_equals_device_path():
    """sfd2_extract_host with one large image uploads it in row bands and runs conv1a band by band behind the copy
    (api.cu: Bands).  Results must equal the device-input path bit for bit, for float32 NCHW and uint8 NHWC, odd heights."""
    from gpu_util import WEIGHTS
    from sfd2_b200 import get_model, extract_resnet_return, Extractor
    from sfd2_b200.synth import synth_image_u8
    model, _ = get_model("ressegnetv2", WEIGHTS, use_stability=True, precision="mixed")
    model.cuda()
    for (H, W) in ((1063, 1600), (1200, 1600), (771, 1029)):
        u8 = synth_image_u8(70 + H, H, W)
        f = torch.from_numpy(np.ascontiguousarray(u8.transpose(2, 0, 1))[None].astype(np.float32) / 255.0)
        a = extract_resnet_return(model, f.pin_memory(), topK=2000, conf_th=0.001)          # host path: banded
        b = extract_resnet_return(model, f.cuda(), topK=2000, conf_th=0.001)                # device path: one conv1a launch
        for k in a:
            assert np.array_equal(a[k], b[k]), (H, W, k)
        ex = Extractor(WEIGHTS, precision="mixed", topk=2000)
        hu = ex.extract_host(torch.from_numpy(u8[None]).pin_memory())
        du = ex(torch.from_numpy(u8[None]).cuda())
        for k in hu:
            assert np.array_equal(hu[k], du[k].cpu().numpy()), (H, W, k, "u8")
