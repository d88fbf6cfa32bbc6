"""The CPU oracle (oracle/sfd2_oracle.py) against fixtures produced by the
UNMODIFIED reference (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import sfd2_oracle as orc
from sfd2_b200.synth import synth_image_u8, shifted_twin


def _img(u8):
    return (u8.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)[None].copy()


def test_generator_is_deterministic(golden):
    g = golden("c1_640x480")
    u8 = synth_image_u8(int(g["seed"]), int(g["H"]), int(g["W"]))
    assert np.array_equal(u8, g["image_u8"])
    g2 = golden("c2_1600x1200")
    u8 = synth_image_u8(int(g2["seed"]), int(g2["H"]), int(g2["W"]))
    assert int(u8.astype(np.int64).sum()) == int(g2["image_sum"])


@pytest.mark.parametrize("name", ["small_96x128", "odd_100x141"])
def test_maps_match_reference(golden, oracle_state, name):
    g = golden(name)
    img = torch.from_numpy(_img(g["image_u8"]))
    x = orc.norm_rgb(img)
    with torch.no_grad():
        score, stab, desc = orc.det(oracle_state, x)
    hm, _ = orc.heatmap(oracle_state, img)
    assert np.array_equal(stab[0, 0].numpy(), g["stability"])
    np.testing.assert_allclose(hm[0, 0].numpy(), g["heat"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(desc[0].numpy(), g["desc_map"], rtol=0, atol=1e-5)
    nms = orc.simple_nms(torch.from_numpy(g["heat"])[None, None], 4)
    assert np.array_equal(nms[0, 0].numpy(), g["nms"])


@pytest.mark.parametrize("name", ["small_96x128", "odd_100x141", "c1_640x480", "c2_1600x1200"])
def test_extract_matches_reference(golden, oracle_state, name):
    g = golden(name)
    H, W, K = int(g["H"]), int(g["W"]), int(g["K"])
    u8 = g["image_u8"] if "image_u8" in g.files else synth_image_u8(int(g["seed"]), H, W)
    out = orc.extract(oracle_state, _img(u8), topK=K, conf_th=0.001, scales=[1.0])
    assert out["keypoints"].dtype == np.float64 and out["descriptors"].dtype == np.float64
    assert np.array_equal(out["keypoints"].astype(np.int16), g["kp_xy"])      # bit-exact indices
    np.testing.assert_allclose(out["scores"], g["scores"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["descriptors"], g["desc"], rtol=0, atol=1e-5)


def test_nms_cases(golden):
    g = golden("nms_cases")
    for k in [f[3:] for f in g.files if f.startswith("in_")]:
        out = orc.simple_nms(torch.from_numpy(g["in_" + k])[None, None], 4)[0, 0].numpy()
        assert np.array_equal(out, g["out_" + k]), k


def test_select_empty_and_single():
    z = torch.zeros(1, 1, 32, 40)
    x, y, s = orc.select_keypoints(z, 0.001, 4, 10)
    assert len(x) == 0
    z[0, 0, 10, 12] = 0.5
    x, y, s = orc.select_keypoints(orc.simple_nms(z, 4), 0.001, 4, 10)
    assert list(x) == [12] and list(y) == [10] and s[0] == 0.5
    z[0, 0, 2, 20] = 0.9   # inside the 4-px border: dropped
    x, y, s = orc.select_keypoints(orc.simple_nms(z, 4), 0.001, 4, 10)
    assert list(x) == [12]


def test_matchers_match_reference(golden):
    g = golden("match_cases")
    for tag in ["sq", "wide", "tall", "one", "col"]:
        d0, d1 = g[f"{tag}_d0"], g[f"{tag}_d1"]
        ph = orc.match_hloc(d0.T[None], d1.T[None])
        assert np.array_equal(ph["matches0"][0].numpy(), g[f"{tag}_hloc_m0"]), tag
        np.testing.assert_allclose(ph["matching_scores0"][0].numpy(), g[f"{tag}_hloc_s0"], atol=1e-6)
        po = orc.match_hloc(d0.T[None], d1.T[None], do_mutual_check=False)
        assert np.array_equal(po["matches0"][0].numpy(), g[f"{tag}_hloc_nomutual_m0"]), tag
        if f"{tag}_itloc_m0" in g.files:
            pi = orc.match_itloc(d0.astype(np.float64), d1.astype(np.float64))
            assert np.array_equal(pi["matches0"], g[f"{tag}_itloc_m0"]), tag
            np.testing.assert_allclose(pi["matching_scores0"], g[f"{tag}_itloc_s0"], atol=1e-12)


def test_pair_matches_reference(golden):
    g = golden("c1_640x480")
    ph = orc.match_hloc(g["desc"].T[None], g["desc_b"].T[None])
    assert np.array_equal(ph["matches0"][0].numpy(), g["hloc_matches0"])
    pi = orc.match_itloc(g["desc"].astype(np.float64), g["desc_b"].astype(np.float64))
    # golden descriptors were stored as float32; the it_loc golden ran on float64
    agree = (pi["matches0"] == g["itloc_matches0"]).mean()
    assert agree > 0.999
    assert (g["hloc_matches0"] == g["itloc_matches0"]).mean() > 0.999


def test_ratio_matchers_match_reference(golden):
    g = golden("match_cases")
    c1 = golden("c1_640x480")
    cases = {t: (g[f"{t}_d0"], g[f"{t}_d1"]) for t in ["sq", "wide", "tall"]}
    cases["c1"] = (c1["desc"], c1["desc_b"])
    for tag, (d0, d1) in cases.items():
        pr = orc.match_hloc(d0.T[None], d1.T[None], ratio_threshold=0.8, distance_threshold=0.7)
        assert np.array_equal(pr["matches0"][0].numpy(), g[f"{tag}_hloc_ratio_m0"]), tag
        np.testing.assert_allclose(pr["matching_scores0"][0].numpy(), g[f"{tag}_hloc_ratio_s0"], atol=1e-6)
        p1 = orc.match_hloc(d0.T[None], d1.T[None], ratio_threshold=0.9, do_mutual_check=False)
        assert np.array_equal(p1["matches0"][0].numpy(), g[f"{tag}_hloc_ratio_nomutual_m0"]), tag
        pi = orc.match_itloc_nnr(d0.astype(np.float64), d1.astype(np.float64), 0.9)
        assert np.array_equal(pi["matches0"], g[f"{tag}_itloc_nnr_m0"]), tag


def test_multiscale_extract_matches_reference(golden, oracle_state):
    g = golden("ms_128x160")
    out = orc.extract(oracle_state, _img(g["image_u8"]), topK=int(g["K"]), scales=list(g["scales"]))
    np.testing.assert_allclose(out["keypoints"], g["kp"], atol=1e-4)
    np.testing.assert_allclose(out["scores"], g["scores"], atol=1e-6)
    np.testing.assert_allclose(out["descriptors"], g["desc"], atol=1e-5)
