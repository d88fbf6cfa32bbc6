"""CPU study for the next round (NOT a test; not collected by pytest): would fp8 correction passes keep the exact
mode's keypoint identity?

The exact mode computes every convolution as  a_hi*w_hi + a_hi*w_lo + a_lo*w_hi  (fp16 hi/lo splits, fp32 accumulate).
The two correction products only need ~5 significant bits; as kind::f8f6f4 MMAs (K = 32 in the cycles of a K = 16
fp16 MMA, profiles/r1s2_mma_rate_probe.log) they would cost half a pass each.  This script emulates the arithmetic
of every mode on the folded network (the same graph the CUDA path runs) with PyTorch on the CPU:

    fp32    : folded fp32 convolutions (the comparison base)
    exact   : the 3-product fp16 split
    fp8corr : main product in fp16, corrections with e4m3 operands and power-of-two scales
    fast    : main product only

and reports, against fp32: shared keypoints of the top-K, relative heat-map error at the candidates, descriptor error.

    python tests/study_fp8_corrections.py [c1|c2]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import sfd2_oracle as orc                       # noqa: E402  (test infrastructure)
from sfd2_b200.weights import fold_layers, load_checkpoint    # noqa: E402

E4M3_MAX = 448.0


def q8(x):
    return x.to(torch.float8_e4m3fn).to(torch.float32)


def pow2_scale(t):
    m = float(t.abs().max())
    return 1.0 if m == 0 else 2.0 ** np.floor(np.log2(E4M3_MAX / m))


def split16(x):
    hi = x.half().float()
    lo = (x - hi).half().float()
    return hi, lo


def conv(x, L, mode, res=None):
    w, b = torch.from_numpy(L["w"]), torch.from_numpy(L["b"])
    k = w.shape[-1]
    kw = dict(stride=L["stride"], padding=k // 2, groups=L["groups"])
    if mode == "fp32":
        y = F.conv2d(x, w, **kw)
    else:
        xh, xl = split16(x)
        wh, wl = split16(w)
        y = F.conv2d(xh, wh, **kw)
        if mode == "exact":
            y = y + (F.conv2d(xh, wl, **kw) + F.conv2d(xl, wh, **kw))
        elif mode == "fp8corr":
            sa, sw = pow2_scale(xh), pow2_scale(wh)
            c = F.conv2d(q8(xh * sa), q8(wl * (sw * 2048.0)), **kw) + F.conv2d(q8(xl * (sa * 2048.0)), q8(wh * sw), **kw)
            y = y + c / (sa * sw * 2048.0)
    y = y + b.view(1, -1, 1, 1)
    if res is not None:
        y = y + res
    if L["relu"]:
        y = F.relu(y)
    if mode != "fp32":              # activations travel as fp16 hi + lo planes
        h, l = split16(y)
        y = h + l
    return y


def forward(layers, img, mode):
    x = orc.norm_rgb(img)
    for n in ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b"]:
        x = conv(x, layers[n], mode)
    for i in range(3):
        t = conv(x, layers[f"rb{i}c1"], mode)
        t = conv(t, layers[f"rb{i}c2"], mode)
        x = conv(t, layers[f"rb{i}c3"], mode, res=x)
    pa = conv(x, layers["convPa0"], mode)
    logits = conv(pa, layers["headP"], mode)          # fp32 out in the CUDA path too (no re-split needed, harmless)
    da = conv(x, layers["convDa0"], mode)
    desc = F.normalize(conv(da, layers["headD"], mode), dim=1)
    sta = conv(x, layers["sta"], "fp32")
    semi = torch.exp(logits[:, :65])
    semi = semi / (semi.sum(1, keepdim=True) + 1e-5)
    score = semi[:, :-1]
    Hc, Wc = score.shape[2:]
    score = score.permute(0, 2, 3, 1).reshape(1, Hc, Wc, 8, 8).permute(0, 1, 3, 2, 4).reshape(1, 1, Hc * 8, Wc * 8)
    stab = orc.cls_to_value(F.interpolate(sta, size=img.shape[2:], mode="bilinear"))
    return score * stab, desc


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c1"
    g = np.load(os.path.join(REPO, "tests", "golden", {"c1": "c1_640x480", "c2": "c2_1600x1200"}[which] + ".npz"))
    K = int(g["K"])
    from sfd2_b200.synth import synth_image
    img = (g["image_u8"].astype(np.float32) / np.float32(255)).transpose(2, 0, 1)[None].copy() if "image_u8" in g.files \
        else synth_image(int(g["seed"]), int(g["H"]), int(g["W"]))
    img = torch.from_numpy(img)
    layers = fold_layers(load_checkpoint(os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz")), prune=False)
    out = {}
    with torch.no_grad():
        for mode in ("fp32", "exact", "fp8corr", "fast"):
            heat, desc = forward(layers, img, mode)
            nms = orc.simple_nms(heat, 4)
            x, y, sc = orc.select_keypoints(nms, 0.001, 4, K)
            d = orc.sample_descriptors(desc, x, y, img.shape[2], img.shape[3])
            out[mode] = (heat, x, y, sc, d)
    h0, x0, y0, s0, d0 = out["fp32"]
    ref = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(x0, y0))}
    gap = abs(float(s0[-1]) - float(g["next_score"])) / float(s0[-1]) if "next_score" in g.files else float("nan")
    print(f"{which}: K = {K}, relative gap between the K-th and (K+1)-th score = {gap:.2e}")
    for mode in ("exact", "fp8corr", "fast"):
        h, x, y, sc, d = out[mode]
        hit = [(i, ref[(int(a), int(b))]) for i, (a, b) in enumerate(zip(x, y)) if (int(a), int(b)) in ref]
        i0, i1 = np.array(hit).T
        rel = (np.abs(h - h0) / h0.clamp_min(1e-12))[0, 0][torch.from_numpy(y0), torch.from_numpy(x0)]
        print(f"  {mode:8s}: {len(hit)}/{K} keypoints shared, heat-map rel. error at the keypoints max {float(rel.max()):.2e} "
              f"median {float(rel.median()):.2e}, descriptor max abs error {np.abs(d[i0] - d0[i1]).max():.2e}")


if __name__ == "__main__":
    main()
