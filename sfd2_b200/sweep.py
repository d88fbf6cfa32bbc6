"""Dataset sweeps over the hot path - the loops the reference runs one item at a time:

  image sweep  <- extract_localization.main:240-272 (every image of a list through the extractor)
  pair sweep   <- extract + hloc.match_features.main:90-121 (extract both frames of a pair, mutual-NN match)

Items are independent, so a multi-GPU sweep is `items[rank::world]` (sfd2_b200.shard) with no collective on
the data path.  Everything stays on the device between the stages: the extractor's fixed-capacity outputs
([n, K, 128] descriptors + counts) are the matcher's inputs, the per-pair row counts are read by the kernels
from device memory, and the host only sees the final tables.
"""
import numpy as np
import torch

from . import _lib
from .extractor import Extractor
from .matchers import match_pairs_dev
from .shard import shard_indices

__all__ = ["Sweep", "image_sweep", "pair_sweep"]


class Sweep:
    """One extractor context + the matcher on one GPU; `batch` images per native extract call."""

    def __init__(self, weight_path, precision="mixed", topk=4096, conf_th=0.001, use_stability=True, device=None,
                 batch=8):
        self.ex = Extractor(weight_path, use_stability=use_stability, precision=precision, topk=topk, conf_th=conf_th,
                            device=device)
        self.precision = precision
        self.topk = int(topk)
        self.batch = int(batch)
        self.device = torch.device("cuda", self.ex.model.ctx.device)

    # ---------------------------------------------------------------- images
    def extract(self, images: torch.Tensor):
        """images: CUDA uint8 [n,H,W,3] or float32 [n,3,H,W] (any n) -> fixed-capacity device tensors."""
        outs = [self.ex(images[i:i + self.batch]) for i in range(0, images.shape[0], self.batch)]
        if len(outs) == 1:
            return outs[0]
        return {k: torch.cat([o[k] for o in outs]) for k in outs[0]}

    def extract_host(self, images):
        """Host images in (pinned uint8 [n,H,W,3] or float32 [n,3,H,W]), host features out; `batch` images per
        sfd2_extract_host call (image i+1's H2D overlaps image i's kernels)."""
        outs = []
        for i in range(0, images.shape[0], self.batch):
            o = self.ex.extract_host(images[i:i + self.batch])
            outs.append({k: v.copy() for k, v in o.items()})      # the pinned result buffers are reused by the next call
        return {k: np.concatenate([o[k] for o in outs]) for k in outs[0]}

    # ---------------------------------------------------------------- pairs
    def match(self, feats, idx0, idx1, mutual=True):
        """Match pair p = (image idx0[p], image idx1[p]) of an extract() result, all pairs in one native call.
        -> matches0 int32 [P, K] (-1 = none; columns beyond counts[idx0[p]] are -1), sim0 float32 [P, K]."""
        return match_pairs_dev(feats["descriptors"], feats["counts"], idx0, idx1, mutual=mutual, precision=self.precision)

    def pairs(self, frames0: torch.Tensor, frames1: torch.Tensor):
        """The C5 unit of work for P pairs: extract frames0[p] and frames1[p], match them.  Device tensors in,
        device tensors out (features of both frames + matches0 / sim0 of frame 0 against frame 1)."""
        P = frames0.shape[0]
        feats = self.extract(torch.cat([frames0, frames1]))
        idx0 = torch.arange(P, device=self.device, dtype=torch.int32)
        m0, s0 = self.match(feats, idx0, idx0 + P)
        return feats, m0, s0

    def pairs_host(self, frames0, frames1):
        """Same with HOST frames (pinned uint8 [P,H,W,3]) in and host numpy out: features of both frames as the
        extract stage stores them, matches0 / sim0 as match_features stores them.  Frames go straight into a persistent
        device buffer (no host-side concatenation), results come back through persistent pinned buffers with one
        synchronisation; the returned arrays are views of those buffers (valid until the next call)."""
        f0, f1 = torch.as_tensor(frames0), torch.as_tensor(frames1)
        P = f0.shape[0]
        shape = (2 * P,) + tuple(f0.shape[1:])
        st = self.__dict__.setdefault("_pairs_stage", {})
        dev = st.get("dev")
        if dev is None or tuple(dev.shape) != shape:
            dev = st["dev"] = torch.empty(shape, dtype=torch.uint8, device=self.device)
        dev[:P].copy_(f0, non_blocking=True)
        dev[P:].copy_(f1, non_blocking=True)
        feats, m0, s0 = self.pairs(dev[:P], dev[P:])
        out = {}
        for k, v in list(feats.items()) + [("matches0", m0), ("sim0", s0)]:
            h = st.get(k)
            if h is None or tuple(h.shape) != tuple(v.shape) or h.dtype != v.dtype:
                h = st[k] = torch.empty(tuple(v.shape), dtype=v.dtype).pin_memory()
            h.copy_(v, non_blocking=True)
            out[k] = h
        self.ex.check_status()                # synchronises the stream: the copies above are complete
        return {k: v.numpy() for k, v in out.items()}


def image_sweep(images_u8, weight_path, rank=0, world=1, **kw):
    """images_u8: sequence of uint8 [H,W,3] arrays (one size).  This rank extracts images[rank::world]; returns
    (indices, {keypoints, scores, descriptors, counts} numpy)."""
    sw = Sweep(weight_path, **kw)
    idx = shard_indices(len(images_u8), rank, world)
    host = torch.from_numpy(np.stack([images_u8[i] for i in idx])).pin_memory()
    return idx, sw.extract_host(host)


def pair_sweep(pairs, weight_path, rank=0, world=1, keep=False, **kw):
    """pairs: sequence of (uint8 [H,W,3], uint8 [H,W,3]).  This rank processes pairs[rank::world].
    -> {"indices", "n_matches" [P]} and, with keep=True, "pairs": per-pair dicts (keypoints0/1, matches0, sim0)."""
    sw = Sweep(weight_path, **kw)
    idx = shard_indices(len(pairs), rank, world)
    out = {"indices": idx, "n_matches": [], "pairs": []}
    for b in range(0, len(idx), max(1, sw.batch // 2)):
        chunk = idx[b:b + max(1, sw.batch // 2)]
        f0 = torch.from_numpy(np.stack([pairs[i][0] for i in chunk])).pin_memory()
        f1 = torch.from_numpy(np.stack([pairs[i][1] for i in chunk])).pin_memory()
        r = sw.pairs_host(f0, f1)
        P = len(chunk)
        for p in range(P):
            n0, n1 = int(r["counts"][p]), int(r["counts"][P + p])
            m0 = r["matches0"][p, :n0]
            out["n_matches"].append(int((m0 >= 0).sum()))
            if keep:
                # copies: r's arrays are views of pinned buffers that the next batch overwrites
                out["pairs"].append({"keypoints0": r["keypoints"][p, :n0].copy(), "keypoints1": r["keypoints"][P + p, :n1].copy(),
                                     "scores0": r["scores"][p, :n0].copy(), "descriptors0": r["descriptors"][p, :n0].copy(),
                                     "descriptors1": r["descriptors"][P + p, :n1].copy(), "matches0": m0.copy(),
                                     "sim0": r["sim0"][p, :n0].copy()})
    return out
