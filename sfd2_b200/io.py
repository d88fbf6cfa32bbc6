"""Feature / match containers and the two per-item loops that sit directly around the hot path.

The reference writes HDF5 (extract_localization.py:235-272, hloc/match_features.py:84-121):
  features:  group <image name> -> keypoints f64[K,2], descriptors f64[128,K], scores f64[K], image_size
  matches :  group names_to_pair(a, b) -> matches0 int16[N], matching_scores0 float16[N]
A path ending in .h5 is a real HDF5 file: written / read with h5py when it is installed, otherwise with the bundled
pure-Python subset writer (h5lite.py: superblock v0, old-style groups, contiguous datasets - what h5py's default writes),
so the files are interchangeable with the reference's.  Any other path is an .npz archive ("<group>/<dataset>" keys).
Descriptors may be stored as float16 (fp16=True): half the bytes of the dominant dataset; readers cast back.
Nothing here computes: extraction and matching go through the CUDA library via extractor.py / matchers.py.
"""
import os

import numpy as np

try:  # optional
    import h5py  # noqa: F401
    HAVE_H5PY = True
except Exception:  # pragma: no cover - h5py is absent in the build image
    HAVE_H5PY = False

__all__ = ["names_to_pair", "Store", "extract_to_store", "match_to_store", "match_confs"]

match_confs = {   # hloc/match_features.py:20-45
    "NNM": {"output": "NNM", "model": {"name": "nearest_neighbor", "do_mutual_check": True, "distance_threshold": None}},
    "ONN": {"output": "ONN", "model": {"name": "nearest_neighbor", "do_mutual_check": False, "distance_threshold": None}},
    "NNR": {"output": "NNR", "model": {"name": "nearest_neighbor", "do_mutual_check": True, "distance_threshold": 0.9}},
}


def names_to_pair(name0, name1):
    """hloc/utils/parsers.py:66-67."""
    return "_".join((name0.replace("/", "-"), name1.replace("/", "-")))


class Store:
    """Group -> {dataset: array} container with the reference's HDF5 layout.
    path ending in .h5 uses h5py (must be installed); anything else is an .npz archive."""

    def __init__(self, path, mode="a", fp16=False, backend=None):
        """fp16: store 'descriptors' as float16 (the reference stores float64: 4.2 MB per 4096 x 128 image -> 1 MB).
        backend: 'h5py' | 'h5lite' | None (h5py when installed)."""
        self.path, self.mode, self.fp16 = str(path), mode, bool(fp16)
        self.h5 = None
        self.groups = {}
        if self.path.endswith(".h5"):
            use_h5py = HAVE_H5PY if backend is None else (backend == "h5py")
            if use_h5py:
                import h5py
                self.h5 = h5py.File(self.path, mode)
            else:
                from . import h5lite
                self.h5 = h5lite.File(self.path, mode)
        elif mode in ("a", "r") and os.path.exists(self.path):
            z = np.load(self.path, allow_pickle=False)
            for key in z.files:
                g, d = key.rsplit("/", 1)
                self.groups.setdefault(g, {})[d] = z[key]

    def __contains__(self, group):
        return (group in self.h5) if self.h5 is not None else (group in self.groups)

    def names(self):
        """Record names (image names / pair keys).  In an HDF5 file a name like 'db/1.jpg' is a nested group: walk down to
        the groups that hold datasets."""
        if self.h5 is None:
            return list(self.groups)
        out = []

        def walk(g, prefix):
            for k in g.keys():
                v = g[k]
                if hasattr(v, "keys"):
                    if any(not hasattr(v[c], "keys") for c in v.keys()):
                        out.append(prefix + k)
                    else:
                        walk(v, prefix + k + "/")
        walk(self.h5, "")
        return out

    def read(self, group):
        if self.h5 is not None:
            return {k: v.__array__() for k, v in self.h5[group].items()}
        return self.groups[group]

    def write(self, group, datasets):
        """create_group semantics: writing an existing group raises, as the reference does
        (extract_localization.py:269 has no skip-if-exists)."""
        if group in self:
            raise ValueError(f"group '{group}' already exists")
        if self.h5 is not None:
            grp = self.h5.create_group(group)
            for k, v in datasets.items():
                grp.create_dataset(k, data=self._cast(k, v))
        else:
            self.groups[group] = {k: self._cast(k, v) for k, v in datasets.items()}

    def _cast(self, key, value):
        a = np.asarray(value)
        return a.astype(np.float16) if (self.fp16 and key == "descriptors") else a

    def close(self):
        if self.h5 is not None:
            self.h5.close()
        elif self.mode != "r":
            np.savez(self.path, **{f"{g}/{d}": a for g, ds in self.groups.items() for d, a in ds.items()})

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def extract_to_store(model, extractor, images, store, conf):
    """The per-image loop of extract_localization.main (:240-272).
    images: iterable of dicts {"name", "image": float tensor [1,3,H,W], "original_size": (w, h)} - what
    ImageDataset yields (:158-190).  conf: the preset's 'model' dict (max_keypoints, conf_th, scales)."""
    n = 0
    for data in images:
        pred = extractor(model, img=data["image"], topK=conf["max_keypoints"], mask=None, conf_th=conf["conf_th"],
                         scales=conf.get("scales", [1.0]))
        pred["descriptors"] = pred["descriptors"].transpose()                        # :253 -> [128, K]
        original_size = np.asarray(data["original_size"])
        pred["image_size"] = original_size
        size = np.array(data["image"].shape[-2:][::-1])
        scales = (original_size / size).astype(np.float32)
        pred["keypoints"] = (pred["keypoints"] + .5) * scales[None] - .5             # :260-263
        store.write(data["name"], pred)
        n += 1
    return n


def match_to_store(model, pair_list, features, store, device="cuda"):
    """The per-pair loop of hloc.match_features.main (:90-121): skip duplicates and pairs already stored,
    descriptors [128,K] -> float32 [1,128,K] on the device, matches0 -> int16, matching_scores0 -> float16."""
    import torch
    matched = set()
    n = 0
    for pair in pair_list:
        name0, name1 = pair.split(" ")
        key = names_to_pair(name0, name1)
        if len({(name0, name1), (name1, name0)} & matched) or key in store:
            continue
        f0, f1 = features.read(name0), features.read(name1)
        data = {}
        for k in f1.keys():
            data[k + "0"] = torch.from_numpy(np.asarray(f0[k]))[None].float().to(device)
            data[k + "1"] = torch.from_numpy(np.asarray(f1[k]))[None].float().to(device)
        pred = model(data)
        out = {"matches0": pred["matches0"][0].cpu().short().numpy()}
        if "matching_scores0" in pred:
            out["matching_scores0"] = pred["matching_scores0"][0].cpu().half().numpy()
        store.write(key, out)
        matched |= {(name0, name1), (name1, name0)}
        n += 1
    return n
