"""The containers and per-item loops around the hot path (extract_localization.main :240-272,
hloc.match_features.main :90-121) with stand-in extract / match callables: layout, dtypes, duplicate
skipping.  (The same loops with the CUDA plugins run in tests/test_gpu_dropin.py.)"""
import numpy as np
import pytest
import torch

from sfd2_b200.io import Store, extract_to_store, match_to_store, names_to_pair


def _fake_extractor(model, img, topK, mask, conf_th, scales):
    rng = np.random.RandomState(int(img.sum() * 1000) % 2**31)
    k = min(topK, 17)
    d = rng.randn(k, 128)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return {"keypoints": rng.rand(k, 2) * 50, "descriptors": d, "scores": np.sort(rng.rand(k))[::-1].copy()}


class _FakeMatcher:
    def __call__(self, data):
        d0, d1 = data["descriptors0"][0], data["descriptors1"][0]          # [128, N], [128, M]
        sim = d0.t() @ d1
        return {"matches0": sim.argmax(1)[None], "matching_scores0": ((sim.max(1).values + 1) / 2)[None]}


def test_names_to_pair():
    assert names_to_pair("db/1.jpg", "query/night/2.jpg") == "db-1.jpg_query-night-2.jpg"


@pytest.mark.parametrize("ext", ["npz", "h5"])
def test_feature_store_layout_and_match_loop(tmp_path, ext):
    """ext = h5: a real HDF5 file (h5py when installed, else the bundled h5lite writer / reader)."""
    imgs = [{"name": f"db/{i}.jpg", "image": torch.full((1, 3, 40, 60), 0.1 * (i + 1)), "original_size": (120, 80)}
            for i in range(3)]
    fpath = tmp_path / f"feats.{ext}"
    with Store(fpath, "w") as st:
        n = extract_to_store(None, _fake_extractor, imgs, st, {"max_keypoints": 12, "conf_th": 0.001, "scales": [1.0]})
        assert n == 3
        with pytest.raises(ValueError):            # the reference's create_group raises on duplicates too
            st.write("db/0.jpg", {"x": np.zeros(1)})
    feats = Store(fpath, "r")
    assert sorted(feats.names()) == ["db/0.jpg", "db/1.jpg", "db/2.jpg"]
    f = feats.read("db/1.jpg")
    assert f["keypoints"].shape == (12, 2) and f["keypoints"].dtype == np.float64
    assert f["descriptors"].shape == (128, 12) and f["scores"].shape == (12,)
    assert list(f["image_size"]) == [120, 80]
    # keypoints were rescaled by original/size = 2 with the half-pixel convention
    ref = _fake_extractor(None, imgs[1]["image"], 12, None, 0.001, [1.0])["keypoints"]
    np.testing.assert_allclose(f["keypoints"], (ref + .5) * 2 - .5)
    pairs = ["db/0.jpg db/1.jpg", "db/1.jpg db/0.jpg", "db/0.jpg db/2.jpg", "db/0.jpg db/1.jpg"]
    mpath = tmp_path / f"matches.{ext}"
    with Store(mpath, "w") as ms:
        assert match_to_store(_FakeMatcher(), pairs, feats, ms, device="cpu") == 2   # reversed + repeated pair skipped
    ms = Store(mpath, "r")
    m = ms.read(names_to_pair("db/0.jpg", "db/2.jpg"))
    assert m["matches0"].dtype == np.int16 and m["matching_scores0"].dtype == np.float16 and m["matches0"].shape == (12,)
    with Store(mpath, "a") as ms2:                  # resume: pairs already in the file are skipped by KEY; the reversed
        # pair "1 0" is a new key and, exactly as in the reference (match_features.py:95-97), is computed now
        assert match_to_store(_FakeMatcher(), pairs + ["db/1.jpg db/2.jpg"], feats, ms2, device="cpu") == 2
