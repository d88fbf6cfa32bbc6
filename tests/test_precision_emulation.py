"""CPU emulation of the precision modes' ARITHMETIC on the folded network (tests/study_fp8_corrections.py): the
design invariants the `mixed` mode rests on, checked without a GPU on the C1 fixture image.

* the fp16 hi/lo 3-product split reproduces the fp32 network's keypoints (1000/1000) with descriptor error < 1e-4;
* running only the descriptor head single-pass leaves keypoints and scores bit-identical (it does not feed the
  heat-map) and keeps the descriptors within the north-star 1e-3.
The CUDA kernels themselves are compared with the oracle in tests/test_gpu_parity.py."""
import os

import numpy as np
import torch

import study_fp8_corrections as S
from oracle import sfd2_oracle as orc
from sfd2_b200.weights import fold_layers, load_checkpoint

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(layers, img, modes, base="exact"):
    orig = S.conv

    def conv(x, L, mode, res=None):
        name = next(k for k, v in layers.items() if v is L)
        return orig(x, L, modes.get(name, mode) if mode != "fp32" else mode, res)
    S.conv = conv
    try:
        with torch.no_grad():
            heat, desc = S.forward(layers, img, base)
    finally:
        S.conv = orig
    x, y, sc = orc.select_keypoints(orc.simple_nms(heat, 4), 0.001, 4, 1000)
    return x, y, sc, orc.sample_descriptors(desc, x, y, img.shape[2], img.shape[3])


def test_mixed_mode_arithmetic_keeps_keypoints_and_bounds_descriptors(golden):
    g = golden("c1_640x480")
    img = torch.from_numpy((g["image_u8"].astype(np.float32) / np.float32(255)).transpose(2, 0, 1)[None].copy())
    layers = fold_layers(load_checkpoint(os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz")), prune=False)
    x0, y0, s0, d0 = _run(layers, img, {}, base="fp32")
    xe, ye, se, de = _run(layers, img, {})
    xm, ym, sm, dm = _run(layers, img, {"convDa0": "fast", "headD": "fast"})
    # exact split == fp32 network on the keypoints, descriptors far inside the tolerance
    assert set(zip(xe.tolist(), ye.tolist())) == set(zip(x0.tolist(), y0.tolist()))
    ref = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(x0, y0))}
    idx = np.array([ref[(int(a), int(b))] for a, b in zip(xe, ye)])
    assert np.abs(de - d0[idx]).max() < 1e-4
    # mixed: the heat-map path is untouched -> identical keypoints and scores, bit for bit
    assert np.array_equal(xm, xe) and np.array_equal(ym, ye) and np.array_equal(sm, se)
    assert 0 < np.abs(dm - de).max() < 1e-3
    assert np.abs(dm - d0[idx]).max() < 1e-3
