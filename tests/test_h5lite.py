"""On-disk formats: the HDF5 layout of the reference's feature / match files
(extract_localization.py:235-272, hloc/match_features.py:84-121), through the bundled pure-Python subset
reader / writer.  The reader is checked on a file written by libhdf5 itself (a MATLAB v7.3 sample that ships with
scipy); the writer is checked through the reader, structurally, and - when h5py is installed - through h5py."""
import os
import struct

import numpy as np
import pytest

from sfd2_b200 import h5lite
from sfd2_b200.io import Store, names_to_pair


def test_reader_on_a_file_written_by_libhdf5():
    import scipy.io
    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy's MATLAB v7.3 sample is not installed")
    with h5lite.File(path, "r") as f:             # 512-byte user block, superblock v0, old-style root group
        assert f.keys() == ["testdouble"]
        a = f["testdouble"].__array__()
        assert a.dtype == np.float64 and a.shape == (9, 1)
        assert np.allclose(a[:, 0], np.arange(9) * np.pi / 4)          # testdouble = 0 : pi/4 : 2 pi


def _features(rng, k):
    return {"keypoints": rng.rand(k, 2) * 1000, "descriptors": rng.randn(128, k), "scores": rng.rand(k),
            "image_size": np.array([1600, 1200])}


def test_round_trip_reference_layout(tmp_path):
    rng = np.random.RandomState(0)
    path = str(tmp_path / "feats.h5")
    names = ["db/1.jpg", "db/2.jpg", "query/day/nexus4/IMG_0001.jpg", "flat.png"]
    ref = {n: _features(rng, 50 + 7 * i) for i, n in enumerate(names)}
    with h5lite.File(path, "w") as f:
        for n, d in ref.items():
            grp = f.create_group(n)                  # nested groups like h5py: 'db' -> '1.jpg'
            for k, v in d.items():
                grp.create_dataset(k, data=v)
        with pytest.raises(ValueError):
            f.create_group("db/1.jpg")               # extract_localization.py:269 relies on this raising
    with h5lite.File(path, "r") as f:
        assert sorted(f.keys()) == ["db", "flat.png", "query"]
        assert "db/2.jpg" in f and "db/3.jpg" not in f
        for n, d in ref.items():
            g = f[n]
            assert sorted(g.keys()) == sorted(d.keys())
            for k, v in d.items():
                got = g[k].__array__()
                assert got.dtype == np.asarray(v).dtype and np.array_equal(got, v), (n, k)
            assert tuple(g["image_size"]) == (1600, 1200)       # match_features.py:108 does tuple(feats['image_size'])
    # append mode keeps what is there and adds to it
    with h5lite.File(path, "a") as f:
        f.create_group("db/3.jpg").create_dataset("scores", data=np.arange(5, dtype=np.float64))
    with h5lite.File(path, "r") as f:
        assert np.array_equal(f["db/3.jpg"]["scores"].__array__(), np.arange(5.0))
        assert np.array_equal(f["db/1.jpg"]["descriptors"].__array__(), ref["db/1.jpg"]["descriptors"])


def test_match_file_dtypes_and_many_groups(tmp_path):
    """matches0 int16 / matching_scores0 float16 (match_features.py:113-119), and enough pairs that the root group's
    B-tree needs a second level (> 2 * INTERNAL_K * 2 * LEAF_K = 8192 entries)."""
    rng = np.random.RandomState(1)
    path = str(tmp_path / "matches.h5")
    n = 8192 + 300
    keys = [names_to_pair(f"db/{i}.jpg", f"query/{(i * 7) % n}.jpg") for i in range(n)]
    with h5lite.File(path, "w") as f:
        for i, k in enumerate(keys):
            g = f.create_group(k)
            g.create_dataset("matches0", data=(rng.randint(-1, 4096, 16)).astype(np.int16))
            g.create_dataset("matching_scores0", data=rng.rand(16).astype(np.float16))
    with h5lite.File(path, "r") as f:
        assert len(f.keys()) == n and set(f.keys()) == set(keys)
        g = f[keys[1234]]
        assert g["matches0"].dtype == np.int16 and g["matching_scores0"].dtype == np.float16
        assert g["matches0"].shape == (16,)
    # structure: superblock v0 with this writer's ranks, root B-tree node at level 1
    raw = open(path, "rb").read()
    assert raw[:8] == h5lite.SIG and raw[8] == 0
    leaf_k, int_k = struct.unpack_from("<HH", raw, 16)
    assert (leaf_k, int_k) == (h5lite.LEAF_K, h5lite.INTERNAL_K)
    eof = struct.unpack_from("<Q", raw, 40)[0]
    assert eof == len(raw)
    bt = struct.unpack_from("<Q", raw, 80)[0]
    assert raw[bt:bt + 4] == b"TREE" and raw[bt + 5] == 1


def test_store_h5_backend_and_fp16_descriptors(tmp_path):
    rng = np.random.RandomState(2)
    path = str(tmp_path / "f.h5")
    feats = _features(rng, 64)
    with Store(path, "a", fp16=True, backend="h5lite") as st:
        st.write("db/a.jpg", feats)
        assert "db/a.jpg" in st
    with Store(path, "r", backend="h5lite") as st:
        r = st.read("db/a.jpg")
        assert np.asarray(r["descriptors"]).dtype == np.float16
        assert np.abs(np.asarray(r["descriptors"], np.float64) - feats["descriptors"]).max() < 2e-3
        assert np.array_equal(np.asarray(r["keypoints"]), feats["keypoints"])


def test_interchange_with_h5py(tmp_path):
    h5py = pytest.importorskip("h5py")
    rng = np.random.RandomState(3)
    feats = _features(rng, 33)
    a, b = str(tmp_path / "ours.h5"), str(tmp_path / "theirs.h5")
    with h5lite.File(a, "w") as f:
        g = f.create_group("db/x.jpg")
        for k, v in feats.items():
            g.create_dataset(k, data=v)
    with h5py.File(a, "r") as f:
        for k, v in feats.items():
            assert np.array_equal(f["db/x.jpg"][k].__array__(), v)
    with h5py.File(b, "w", libver="earliest") as f:
        g = f.create_group("db/x.jpg")
        for k, v in feats.items():
            g.create_dataset(k, data=v)
    with h5lite.File(b, "r") as f:
        for k, v in feats.items():
            assert np.array_equal(f["db/x.jpg"][k].__array__(), v)


def test_rejects_what_it_does_not_implement(tmp_path):
    with pytest.raises(h5lite.H5Error):
        with h5lite.File(str(tmp_path / "s.h5"), "w") as f:
            f.create_dataset("names", data=np.array(["a", "b"]))
    p = tmp_path / "junk.h5"
    p.write_bytes(b"not hdf5" * 100)
    with pytest.raises(h5lite.H5Error):
        h5lite.File(str(p), "r")
