"""Host-side logic (no GPU): weight folding / packing against the oracle network,
interface shapes of the drop-ins, synthetic generators."""
import struct

import numpy as np
import torch
import torch.nn.functional as F

from oracle import sfd2_oracle as orc
from sfd2_b200 import weights as wts
from sfd2_b200.synth import synth_image, synth_descriptors


def _run_folded(layers, x):
    """The folded + merged network evaluated with plain F.conv2d (float64): must equal the
    oracle's unfolded conv->BN->ReLU network up to rounding."""
    def c(name, t, res=None):
        L = layers[name]
        y = F.conv2d(t, torch.from_numpy(L["w"]).double(), torch.from_numpy(L["b"]).double(), stride=L["stride"],
                     padding=L["w"].shape[2] // 2, groups=L["groups"])
        if res is not None:
            y = y + res
        return F.relu(y) if L["relu"] else y
    t = x.double()
    for n in ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b"]:
        t = c(n, t)
    for i in range(3):
        t = c(f"rb{i}c3", c(f"rb{i}c2", c(f"rb{i}c1", t)), res=t)
    logits = c("headP", c("convPa0", t))
    desc = c("headD", c("convDa0", t))
    sta = c("sta", t)
    return t, logits, desc, sta


def test_fold_and_merge_equal_oracle(oracle_state):
    layers = wts.fold_layers({k: v.numpy() for k, v in oracle_state.items()})
    assert list(layers) == wts.LAYER_ORDER
    x = orc.norm_rgb(torch.from_numpy(synth_image(2, 64, 96)))
    out4, logits, desc, sta = _run_folded(layers, x)
    with torch.no_grad():
        ref4 = orc.backbone(oracle_state, x)
        score, stab, rdesc = orc.det(oracle_state, x)
    np.testing.assert_allclose(out4.numpy(), ref4.numpy(), atol=2e-4)
    semi = torch.exp(logits)
    semi = semi / (semi.sum(1, keepdim=True) + 1e-5)
    sc = semi[:, :-1].permute(0, 2, 3, 1).reshape(1, 8, 12, 8, 8).permute(0, 1, 3, 2, 4).reshape(1, 1, 64, 96)
    np.testing.assert_allclose(sc.numpy(), score.numpy(), atol=2e-6)
    np.testing.assert_allclose(F.normalize(desc, dim=1).numpy(), rdesc.numpy(), atol=2e-5)


def test_blob_layout(oracle_state):
    layers = wts.fold_layers({k: v.numpy() for k, v in oracle_state.items()})
    blob = wts.pack_blob(layers)
    assert blob[:8] == b"SFD2W001"
    n, = struct.unpack_from("<I", blob, 8)
    assert n == len(wts.LAYER_ORDER)
    name, cin, cout, k, stride, groups, relu, woff, boff = struct.unpack_from("<16s6i2Q", blob, 16 + 7 * 56)
    assert name.rstrip(b"\0") == b"rb0c2" and (cin, cout, k, stride, groups, relu) == (256, 256, 3, 1, 32, 1)
    w = np.frombuffer(blob, np.float32, 256 * 8 * 9, woff).reshape(256, 8, 3, 3)
    assert np.array_equal(w, layers["rb0c2"]["w"])
    assert boff == woff + w.nbytes


def test_synth_descriptors_are_unit_norm():
    d0, d1 = synth_descriptors(0, 100, 60)
    np.testing.assert_allclose(np.linalg.norm(d0, axis=1), 1, atol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(d1, axis=1), 1, atol=1e-6)
