"""Checkpoint -> folded weight blob for libsfd2_b200.so.

Takes the reference's own state dict (the 'model' entry of the .pth that
extract_localization.get_model loads, extract_localization.py:214-215, or the
.npz export of it) and applies the algebraic rewrites that only change rounding
(SURVEY.md A.1): every eval-mode BatchNorm is folded into the preceding conv in
float64, and the two head pairs without a non-linearity between them
(convPb o convPa.3, convDb o convDa.3; nets/sfd2.py:286-300, :328-341) are
pre-multiplied into one 3x3 conv each.  The result is rounded to fp32 once.
"""
import struct

import numpy as np

BN_EPS = 1e-5
MAGIC = b"SFD2W001"

LAYER_ORDER = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b",
               "rb0c1", "rb0c2", "rb0c3", "rb1c1", "rb1c2", "rb1c3", "rb2c1", "rb2c2", "rb2c3",
               "convPa0", "headP", "convDa0", "headD", "sta"]


def load_checkpoint(path: str) -> dict:
    """name -> float64 ndarray, from a reference .pth or an .npz export."""
    if str(path).endswith(".npz"):
        z = np.load(path)
        return {k: np.asarray(z[k], dtype=np.float64) for k in z.files}
    import torch
    ck = torch.load(path, map_location="cpu", weights_only=False)
    sd = ck["model"] if isinstance(ck, dict) and "model" in ck else ck
    return {k: v.double().numpy() for k, v in sd.items() if hasattr(v, "numpy") and v.ndim > 0}


def _fold(w, b, sd, bn, affine):
    scale = 1.0 / np.sqrt(sd[bn + ".running_var"] + BN_EPS)
    shift = -sd[bn + ".running_mean"] * scale
    if affine:
        g, be = sd[bn + ".weight"], sd[bn + ".bias"]
        scale, shift = scale * g, shift * g + be
    if b is None:
        b = np.zeros(w.shape[0])
    return w * scale[:, None, None, None], b * scale + shift


def fold_layers(sd: dict, prune: bool = True) -> dict:
    """-> {name: dict(w=f32 OIHW, b=f32, stride, groups, relu)} in LAYER_ORDER."""
    sd = {k: np.asarray(v, dtype=np.float64) for k, v in sd.items()}
    out = {}

    def put(name, w, b, stride=1, groups=1, relu=1):
        out[name] = dict(w=np.ascontiguousarray(w, dtype=np.float32), b=np.ascontiguousarray(b, dtype=np.float32),
                         stride=stride, groups=groups, relu=relu)

    for name, conv, bn, stride in [("conv1a", "conv1a.0", "conv1a.1", 1), ("conv1b", "conv1b.0", "bn1b.0", 2),
                                   ("conv2a", "conv2a.0", "conv2a.1", 1), ("conv2b", "conv2b.0", "bn2b.0", 2),
                                   ("conv3a", "conv3a.0", "conv3a.1", 1), ("conv3b", "conv3b.0", "bn3b.0", 1)]:
        w, b = _fold(sd[conv + ".weight"], sd[conv + ".bias"], sd, bn, affine=False)
        put(name, w, b, stride=stride)
    for i in range(3):
        p = f"conv4.{i}"
        w, b = _fold(sd[p + ".conv1.weight"], None, sd, p + ".bn1", True)
        put(f"rb{i}c1", w, b)
        w, b = _fold(sd[p + ".conv2.weight"], None, sd, p + ".bn2", True)
        put(f"rb{i}c2", w, b, groups=32)
        w, b = _fold(sd[p + ".conv3.weight"], None, sd, p + ".bn3", True)
        put(f"rb{i}c3", w, b, relu=1)       # ReLU after the residual add (nets/sfd2.py:52-53)
    for head, a0, bn, a3, bb in [("P", "convPa.0", "convPa.1", "convPa.3", "convPb"),
                                 ("D", "convDa.0", "convDa.1", "convDa.3", "convDb")]:
        w, b = _fold(sd[a0 + ".weight"], sd[a0 + ".bias"], sd, bn, True)
        put(f"conv{head}a0", w, b, stride=2 if head == "P" else 1)
        w3, b3 = sd[a3 + ".weight"], sd[a3 + ".bias"]               # [256,256,3,3]
        wb, bbias = sd[bb + ".weight"][:, :, 0, 0], sd[bb + ".bias"]  # [o,256]
        wm = np.einsum("om,mikl->oikl", wb, w3)
        bm = wb @ b3 + bbias
        put(f"head{head}", wm, bm, relu=0)
    if "ConvSta.weight" in sd:
        put("sta", sd["ConvSta.weight"], sd["ConvSta.bias"], relu=0)
    else:   # require_stability=False checkpoints: a zero head (never evaluated)
        put("sta", np.zeros((3, 256, 1, 1)), np.zeros(3), relu=0)
    if prune:
        _prune_dead(out, "convPa0", "headP")
        _prune_dead(out, "convDa0", "headD")
    return {k: out[k] for k in LAYER_ORDER}


DEAD_EPS = 1e-20


def _prune_dead(layers, producer, consumer):
    """Drop the producer's dead output channels and the consumer's matching input channels.

    The checkpoint has BatchNorm channels with running_var ~ 5.6e-45 and gamma ~ 1e-40 (SURVEY.md
    item 9): after folding, every weight of such a channel is < 1e-36 and its bias is < 2e-28, so
    the channel is the constant relu(bias) ~ 0 and contributes < 1e-27 to the next layer - far below
    fp32 resolution of the logits it feeds.  Removing it changes nothing representable in fp32 and
    removes 133/256 (convPa.0) and 70/256 (convDa.0) of those layers' work.  Live channels are padded
    with all-zero channels up to a multiple of 64 (the tensor-core K chunk)."""
    P, Cn = layers[producer], layers[consumer]
    w, b = P["w"], P["b"]
    live = (np.abs(w).reshape(w.shape[0], -1).max(1) >= DEAD_EPS) | (np.maximum(b, 0) >= DEAD_EPS)
    idx = np.nonzero(live)[0]
    n_keep = -(-len(idx) // 64) * 64
    if n_keep >= w.shape[0]:
        return
    wp = np.zeros((n_keep,) + w.shape[1:], np.float32)
    bp = np.zeros((n_keep,), np.float32)
    wp[:len(idx)], bp[:len(idx)] = w[idx], b[idx]
    P["w"], P["b"] = wp, bp
    wc = Cn["w"]
    wn = np.zeros((wc.shape[0], n_keep) + wc.shape[2:], np.float32)
    wn[:, :len(idx)] = wc[:, idx]
    Cn["w"] = wn


def pack_blob(layers: dict) -> bytes:
    """Serialise folded layers into the blob sfd2_create parses (csrc/api.cu)."""
    n = len(layers)
    head = 16 + n * 56
    table, data, off = [], [], head
    for name, L in layers.items():
        w, b = L["w"], L["b"]
        cout, cpg, k, _ = w.shape
        wb, bb = w.tobytes(), b.tobytes()
        table.append(struct.pack("<16s6i2Q", name.encode(), cpg * L["groups"], cout, k, L["stride"], L["groups"],
                                 L["relu"], off, off + len(wb)))
        data += [wb, bb]
        off += len(wb) + len(bb)
    return MAGIC + struct.pack("<II", n, 0) + b"".join(table) + b"".join(data)


def blob_from_checkpoint(path: str) -> bytes:
    return pack_blob(fold_layers(load_checkpoint(path)))
