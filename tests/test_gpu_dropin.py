"""Drop-in level: the reference's two CLI loops (extract -> feature store -> pair matching -> match store) with
the CUDA plugins, against the same loops run with the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import sfd2_oracle as orc
from sfd2_b200.io import Store, extract_to_store, match_to_store, names_to_pair, match_confs
from sfd2_b200.synth import synth_image_u8, shifted_twin

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_extract_then_match_pipeline(tmp_path, oracle_state):
    from gpu_util import WEIGHTS
    from sfd2_b200 import get_model, NearestNeighbor
    u8 = synth_image_u8(4, 240, 320)
    frames = {"a.jpg": u8, "b.jpg": shifted_twin(u8), "c.jpg": synth_image_u8(5, 240, 320)}
    imgs = [{"name": k, "image": torch.from_numpy((v.astype(np.float32) / np.float32(255)).transpose(2, 0, 1)[None].copy()),
             "original_size": (640, 480)} for k, v in frames.items()]
    conf = {"max_keypoints": 400, "conf_th": 0.001, "scales": [1.0]}
    model, extractor = get_model("ressegnetv2", WEIGHTS, use_stability=True)
    model = model.cuda()
    with Store(tmp_path / "f.npz", "w") as st:
        extract_to_store(model, extractor, imgs, st, conf)
    with Store(tmp_path / "f_ref.npz", "w") as st:
        extract_to_store(oracle_state, lambda m, img, topK, mask, conf_th, scales: orc.extract(m, img, topK, conf_th, scales),
                         imgs, st, conf)
    ours, ref = Store(tmp_path / "f.npz", "r"), Store(tmp_path / "f_ref.npz", "r")
    for name in frames:
        a, b = ours.read(name), ref.read(name)
        n = b["keypoints"].shape[0]
        assert 100 < n <= 400 and a["descriptors"].shape == (128, n) and a["keypoints"].dtype == np.float64
        ka, kb = set(map(tuple, np.round(a["keypoints"], 3))), set(map(tuple, np.round(b["keypoints"], 3)))
        assert len(ka & kb) >= n - 2
    pairs = ["a.jpg b.jpg", "a.jpg c.jpg", "b.jpg a.jpg"]
    nn = NearestNeighbor(match_confs["NNM"]["model"]).eval().to("cuda")
    with Store(tmp_path / "m.npz", "w") as ms:
        assert match_to_store(nn, pairs, ours, ms, device="cuda") == 2
    ms = Store(tmp_path / "m.npz", "r")
    m_ab = ms.read(names_to_pair("a.jpg", "b.jpg"))
    assert m_ab["matches0"].dtype == np.int16 and m_ab["matching_scores0"].dtype == np.float16
    fa, fb = ours.read("a.jpg"), ours.read("b.jpg")
    refm = orc.match_hloc(fa["descriptors"][None], fb["descriptors"][None])["matches0"][0].numpy()
    assert (m_ab["matches0"] == refm).mean() > 0.995
    assert (m_ab["matches0"] >= 0).sum() > 100          # the shifted twin matches; the unrelated frame does not
    assert (ms.read(names_to_pair("a.jpg", "c.jpg"))["matches0"] >= 0).sum() < (m_ab["matches0"] >= 0).sum()
