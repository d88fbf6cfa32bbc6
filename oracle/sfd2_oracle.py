"""CPU ORACLE for the SFD2 extract + match hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain PyTorch-CPU / numpy restatement of the reference algorithm
(feixue94/sfd2 @ f37fe0c).  It is the checker the CUDA path is compared with and
the CPU baseline `bench.py` times; it is never imported by the product package
(`sfd2_b200/`), which fails loudly when its CUDA library is missing.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it.

Pinning: the reference has no tests or golden vectors of its own (SURVEY.md §4,
§8c).  The oracle is pinned against the reference ITSELF, imported from
/root/reference in the build container by `oracle/make_golden.py`, which writes
`tests/golden/*.npz`; `tests/test_oracle.py` checks this restatement against
those fixtures on every run (bit-exact for NMS / selection / matches, 1e-5 for
floating-point maps).

Every function cites the reference lines it follows.  The op sequence is kept
literally the same as the reference's (unfused conv -> BN -> ReLU, five
max_pool2d calls in NMS, grid_sample, a second GEMM in the it_loc matcher) so
that timing it is a fair stand-in for the reference's own CPU path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

RGB_MEAN = (0.485, 0.456, 0.406)   # nets/extractor.py:14
RGB_STD = (0.229, 0.224, 0.225)    # nets/extractor.py:15
BN_EPS = 1e-5                      # nn.BatchNorm2d default, nets/sfd2.py:58-65,85


# --------------------------------------------------------------------------- weights
def load_state(path: str) -> dict:
    """Load the checkpoint tensors exported by oracle/export_weights.py (npz with
    the reference's own state-dict key names, fp32)."""
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) for k in z.files}


# --------------------------------------------------------------------------- network
def _bn(st, name, x, affine):
    """Eval-mode BatchNorm2d (running stats), nets/sfd2.py:58-65 / :85 / :31."""
    w = st[name + ".weight"] if affine else None
    b = st[name + ".bias"] if affine else None
    return F.batch_norm(x, st[name + ".running_mean"], st[name + ".running_var"], w, b,
                        training=False, eps=BN_EPS)


def _conv(st, name, x, stride=1, padding=1, groups=1):
    return F.conv2d(x, st[name + ".weight"], st.get(name + ".bias"), stride=stride,
                    padding=padding, groups=groups)


def _resblock(st, p, x):
    """ResBlock.forward, nets/sfd2.py:38-55 (1x1 -> grouped 3x3 g=32 -> 1x1, +id)."""
    out = F.relu(_bn(st, p + ".bn1", _conv(st, p + ".conv1", x, padding=0), True))
    out = F.relu(_bn(st, p + ".bn2", _conv(st, p + ".conv2", out, groups=32), True))
    out = _bn(st, p + ".bn3", _conv(st, p + ".conv3", out, padding=0), True)
    return F.relu(out + x)


def backbone(st, x):
    """conv1a..conv4 of ResSegNetV2.det, nets/sfd2.py:314-326 (layers :268-284)."""
    x = F.relu(_bn(st, "conv1a.1", _conv(st, "conv1a.0", x), False))
    x = F.relu(_bn(st, "bn1b.0", _conv(st, "conv1b.0", x, stride=2), False))
    x = F.relu(_bn(st, "conv2a.1", _conv(st, "conv2a.0", x), False))
    x = F.relu(_bn(st, "bn2b.0", _conv(st, "conv2b.0", x, stride=2), False))
    x = F.relu(_bn(st, "conv3a.1", _conv(st, "conv3a.0", x), False))
    x = F.relu(_bn(st, "bn3b.0", _conv(st, "conv3b.0", x), False))
    for i in range(3):
        x = _resblock(st, f"conv4.{i}", x)
    return x


def cls_to_value(x):
    """nets/sfd2.py:305-311: argmax over 3 classes -> {0:0.1, 1:0.5, 2:1.0}."""
    cls = torch.max(x, dim=1, keepdim=True)[1]
    stab = torch.ones_like(cls).float()
    stab[cls == 0] = 0.1
    stab[cls == 1] = 0.5
    return stab


def det(st, x, require_stability=True):
    """ResSegNetV2.det, nets/sfd2.py:313-354 -> (score[1,1,H8*8,W8*8], stability, desc)."""
    out4 = backbone(st, x)
    cPa = F.relu(_bn(st, "convPa.1", _conv(st, "convPa.0", out4, stride=2), True))
    cPa = _conv(st, "convPa.3", cPa)
    semi = _conv(st, "convPb", cPa, padding=0)
    semi = torch.exp(semi)                                              # :330
    semi_norm = semi / (torch.sum(semi, dim=1, keepdim=True) + .00001)  # :331
    score = semi_norm[:, :-1, :, :]
    Hc, Wc = score.size(2), score.size(3)
    score = score.permute([0, 2, 3, 1]).view(score.size(0), Hc, Wc, 8, 8)
    score = score.permute([0, 1, 3, 2, 4]).contiguous().view(score.size(0), 1, Hc * 8, Wc * 8)

    cDa = F.relu(_bn(st, "convDa.1", _conv(st, "convDa.0", out4), True))
    cDa = _conv(st, "convDa.3", cDa)
    desc = F.normalize(_conv(st, "convDb", cDa, padding=0), dim=1)      # :340-342

    stability = None
    if require_stability:
        stability = _conv(st, "ConvSta", out4, padding=0)               # :345
        stability = F.interpolate(stability, size=(x.shape[2], x.shape[3]), mode="bilinear")
        stability = cls_to_value(stability)
    return score, stability, desc


# --------------------------------------------------------------------------- post-processing
def norm_rgb(img):
    """tvf.Normalize(mean, std), nets/extractor.py:14-17."""
    mean = torch.tensor(RGB_MEAN, dtype=img.dtype).view(1, 3, 1, 1)
    std = torch.tensor(RGB_STD, dtype=img.dtype).view(1, 3, 1, 1)
    return (img - mean) / std


def simple_nms(scores, nms_radius: int):
    """nets/extractor.py:20-35, literally (5 max-pools, 2 refinement rounds)."""
    def max_pool(x):
        return F.max_pool2d(x, kernel_size=nms_radius * 2 + 1, stride=1, padding=nms_radius)

    zeros = torch.zeros_like(scores)
    max_mask = scores == max_pool(scores)
    for _ in range(2):
        supp_mask = max_pool(max_mask.float()) > 0
        supp_scores = torch.where(supp_mask, zeros, scores)
        new_max_mask = supp_scores == max_pool(supp_scores)
        max_mask = max_mask | (new_max_mask & (~supp_mask))
    return torch.where(max_mask, scores, zeros)


def heatmap(st, img, use_stability=True):
    """nets/extractor.py:104-141: normalise, det, (resize), x stability."""
    x = norm_rgb(img.reshape(1, *img.shape[-3:]).float())
    nh, nw = x.shape[2:]
    with torch.no_grad():
        hm, stab, desc = det(st, x, require_stability=use_stability)
        if hm.size(2) != nh or hm.size(3) != nw:
            hm = F.interpolate(hm, size=[nh, nw], mode="bilinear", align_corners=False)  # :137-138
        if stab is not None:
            hm = hm * stab
    return hm, desc


def canonical_order(scores: np.ndarray, x: np.ndarray, y: np.ndarray, width: int) -> np.ndarray:
    """The deterministic order both sides are compared in (the reference's own is
    np.argsort-unstable, nets/extractor.py:176,323): score descending, then
    y*W+x ascending."""
    lin = y.astype(np.int64) * width + x.astype(np.int64)
    return np.lexsort((lin, -scores.astype(np.float64)))


def select_keypoints(nms: torch.Tensor, conf_th: float, border: int, topK: int):
    """nets/extractor.py:158-183 + :322-326 for one scale: threshold, border
    removal, sort by score, top-K.  Returns x[int64], y[int64], score[float32]
    in canonical order (ties by pixel index)."""
    s = nms.reshape(nms.shape[-2], nms.shape[-1])
    H, W = s.shape
    kp = torch.nonzero(s > conf_th)                 # (y, x), row-major
    sc = s[kp[:, 0], kp[:, 1]].numpy()
    y = kp[:, 0].numpy()
    x = kp[:, 1].numpy()
    keep = ~((x < border) | (x >= W - border) | (y < border) | (y >= H - border))
    x, y, sc = x[keep], y[keep], sc[keep]
    order = canonical_order(sc, x, y, W)
    if topK > 0:
        order = order[:topK]
    return x[order], y[order], sc[order]


def sample_descriptors(desc: torch.Tensor, x: np.ndarray, y: np.ndarray, nh: int, nw: int) -> np.ndarray:
    """nets/extractor.py:199-208: grid_sample (bilinear, zeros, align_corners=False)
    at (x/(nw/2)-1, y/(nh/2)-1), then divide by the L2 norm.  -> [n, D] float32."""
    D = desc.size(1)
    if len(x) == 0:
        return np.zeros((0, D), np.float32)
    samp = torch.from_numpy(np.stack([x, y]).astype(np.float32))   # pts are float32 in the reference (:163)
    samp[0, :] = (samp[0, :] / (float(nw) / 2.)) - 1.
    samp[1, :] = (samp[1, :] / (float(nh) / 2.)) - 1.
    samp = samp.transpose(0, 1).contiguous().view(1, 1, -1, 2).float()
    d = F.grid_sample(desc, samp, mode="bilinear", padding_mode="zeros", align_corners=False)
    d = d.numpy().reshape(D, -1)
    d = d / np.linalg.norm(d, axis=0)[np.newaxis, :]
    return np.ascontiguousarray(d.T)


def extract(st, img, topK=-1, conf_th=0.001, scales=(1.0,), use_stability=True):
    """extract_resnet_return(model, img, conf_th, mask=None, topK, scales=...),
    nets/extractor.py:97-337, single- or multi-scale, mask=None branch.
    img: float tensor/array [1,3,H,W] (or [3,H,W]) in [0,1], RGB.
    Returns {"keypoints": f64[K,2], "descriptors": f64[K,128], "scores": f64[K]}
    in canonical order.  Unlike the reference (which crashes on 0 keypoints,
    :219), empty inputs give empty arrays."""
    img = torch.as_tensor(np.asarray(img)).float()
    img = img.reshape(1, *img.shape[-3:])
    _, _, H, W = img.shape
    pts_all, desc_all, lin_all = [], [], []
    for s in scales:
        if s == 1.0:
            new_img = img
        else:   # :122-125 -- the reference resizes the *normalised* image; the two commute
            new_img = F.interpolate(img, size=(int(H * s), int(W * s)), mode="bilinear",
                                    align_corners=False)
        nh, nw = new_img.shape[2:]
        if s == 1.0:
            hm, desc = heatmap(st, new_img, use_stability)
        else:
            x = F.interpolate(norm_rgb(img), size=(nh, nw), mode="bilinear", align_corners=False)
            with torch.no_grad():
                hm, stab, desc = det(st, x, require_stability=use_stability)
                if hm.size(2) != nh or hm.size(3) != nw:
                    hm = F.interpolate(hm, size=[nh, nw], mode="bilinear", align_corners=False)
                if stab is not None:
                    hm = hm * stab
        nms = simple_nms(hm, 4)
        # border test uses the ORIGINAL W/H (:181-182), not nw/nh
        sgrid = nms.reshape(nh, nw)
        kp = torch.nonzero(sgrid > conf_th)
        sc = sgrid[kp[:, 0], kp[:, 1]].numpy()
        y = kp[:, 0].numpy()
        x = kp[:, 1].numpy()
        keep = ~((x < 4) | (x >= W - 4) | (y < 4) | (y >= H - 4))
        x, y, sc = x[keep], y[keep], sc[keep]
        d = sample_descriptors(desc, x, y, nh, nw) if (desc.size(2) != nh or desc.size(3) != nw) \
            else desc[0][:, y, x].numpy().T
        if len(x) == 0:
            continue
        pts = np.stack([x.astype(np.float32) * W / nw, y.astype(np.float32) * H / nh, sc], 1)
        pts_all.append(pts.astype(np.float32))
        desc_all.append(d)
        lin_all.append(y.astype(np.int64) * nw + x.astype(np.int64))
    if not pts_all:
        return {"keypoints": np.zeros((0, 2)), "descriptors": np.zeros((0, 128)), "scores": np.zeros((0,))}
    pts = np.vstack(pts_all)
    descs = np.vstack(desc_all)
    lin = np.concatenate(lin_all)
    order = np.lexsort((lin, -pts[:, 2].astype(np.float64)))
    if topK > 0:
        order = order[:topK]
    return {"keypoints": np.array(pts[order, 0:2], dtype=float),
            "descriptors": np.array(descs[order], dtype=float),
            "scores": np.array(pts[order, 2], dtype=float)}


# --------------------------------------------------------------------------- matchers
def find_nn(sim, ratio_thresh, distance_thresh):
    """hloc/matchers/nearest_neighbor.py:6-16."""
    sim_nn, ind_nn = sim.topk(2 if ratio_thresh else 1, dim=-1, largest=True)
    dist_nn = 2 * (1 - sim_nn)
    mask = torch.ones(ind_nn.shape[:-1], dtype=torch.bool)
    if ratio_thresh:
        mask = mask & (dist_nn[..., 0] <= (ratio_thresh ** 2) * dist_nn[..., 1])
    if distance_thresh:
        mask = mask & (dist_nn[..., 0] <= distance_thresh ** 2)
    matches = torch.where(mask, ind_nn[..., 0], ind_nn.new_tensor(-1))
    scores = torch.where(mask, (sim_nn[..., 0] + 1) / 2, sim_nn.new_tensor(0))
    return matches, scores


def mutual_check(m0, m1):
    """hloc/matchers/nearest_neighbor.py:19-24."""
    inds0 = torch.arange(m0.shape[-1])
    loop = torch.gather(m1, -1, torch.where(m0 > -1, m0, m0.new_tensor(0)))
    ok = (m0 > -1) & (inds0 == loop)
    return torch.where(ok, m0, m0.new_tensor(-1))


def match_hloc(desc0, desc1, ratio_threshold=None, distance_threshold=None, do_mutual_check=True):
    """NearestNeighbor._forward, hloc/matchers/nearest_neighbor.py:38-57.
    desc0 [1,D,N], desc1 [1,D,M] float32 -> matches0 int64 [1,N], matching_scores0 f32 [1,N]."""
    desc0 = torch.as_tensor(desc0).float()
    desc1 = torch.as_tensor(desc1).float()
    sim = torch.einsum("bdn,bdm->bnm", desc0, desc1)
    matches0, scores0 = find_nn(sim, ratio_threshold, distance_threshold)
    if do_mutual_check:
        matches1, _ = find_nn(sim.transpose(1, 2), ratio_threshold, distance_threshold)
        matches0 = mutual_check(matches0, matches1)
    return {"matches0": matches0, "matching_scores0": scores0}


def match_itloc(descriptors0: np.ndarray, descriptors1: np.ndarray):
    """it_loc/matcher.py Matcher.forward (mode 'nnm') :91-119 + mutual_nn_matcher
    :122-130: numpy [N,D], [M,D] in the caller's dtype (float64 from h5) ->
    matches0 int[N] (-1 = none), matching_scores0 = raw max cosine per row."""
    d1 = torch.from_numpy(np.ascontiguousarray(descriptors0))
    d2 = torch.from_numpy(np.ascontiguousarray(descriptors1))
    sim = d1 @ d2.t()
    nn12 = torch.max(sim, dim=1)[1]
    nn21 = torch.max(sim, dim=0)[1]
    ids1 = torch.arange(0, sim.shape[0])
    mask = ids1 == nn21[nn12]
    matches = torch.stack([ids1[mask], nn12[mask]]).t().numpy()
    all_matches = np.ones((d1.shape[0],), dtype=int) * -1
    scores = torch.topk(d1 @ d2.t(), dim=1, k=1)[0]          # the reference's second GEMM (:113)
    for i in range(matches.shape[0]):
        all_matches[matches[i, 0]] = matches[i, 1]
    return {"matches0": all_matches, "matching_scores0": scores.squeeze(-1).numpy()}


def match_itloc_nnr(descriptors0: np.ndarray, descriptors1: np.ndarray, ratio: float = 0.9):
    """it_loc/matcher.py mode 'nnr' (:101-103) -> mutual_nn_ratio_matcher (:165-194): mutual NN plus the
    symmetric Lowe ratio test sqrt(2-2 s0) / (sqrt(2-2 s1) + 1e-8) <= ratio in both directions."""
    d1 = torch.from_numpy(np.ascontiguousarray(descriptors0))
    d2 = torch.from_numpy(np.ascontiguousarray(descriptors1))
    sim = d1 @ d2.t()
    nns_sim, nns = torch.topk(sim, 2, dim=1)
    nns_dist = torch.sqrt(2 - 2 * nns_sim)
    ratios12 = nns_dist[:, 0] / (nns_dist[:, 1] + 1e-8)
    nn12 = nns[:, 0]
    nns_sim, nns = torch.topk(sim.t(), 2, dim=1)
    nns_dist = torch.sqrt(2 - 2 * nns_sim)
    ratios21 = nns_dist[:, 0] / (nns_dist[:, 1] + 1e-8)
    nn21 = nns[:, 0]
    ids1 = torch.arange(0, sim.shape[0])
    mask = torch.min(ids1 == nn21[nn12], torch.min(ratios12 <= ratio, ratios21[nn12] <= ratio))
    all_matches = np.ones((d1.shape[0],), dtype=int) * -1
    all_matches[ids1[mask].numpy()] = nn12[mask].numpy()
    scores = torch.topk(sim, dim=1, k=1)[0]
    return {"matches0": all_matches, "matching_scores0": scores.squeeze(-1).numpy()}


def feature_matching(desc_q: np.ndarray, desc_db: np.ndarray, db_3D_ids=None):
    """it_loc/localize_cv2.py:511-560 without labels, matcher = Matcher(confs['NNM']) (it_loc/matcher.py:85-130):
    only db keypoints with a 3-D point take part (desc_db[db_3D_ids != -1]); matches are mapped back to the
    original db rows (:557-559); <= 3 valid db keypoints -> no matches (:537-538)."""
    if db_3D_ids is None:
        return match_itloc(desc_q, desc_db)["matches0"]
    masks = (np.asarray(db_3D_ids) != -1)
    if np.sum(masks) <= 3:
        return np.ones((desc_q.shape[0],), dtype=int) * -1
    valid_ids = np.nonzero(masks)[0]
    matches = match_itloc(desc_q, desc_db[masks])["matches0"].copy()
    ok = matches >= 0
    matches[ok] = valid_ids[matches[ok]]
    return matches


def mutual_nn_exact(d0: np.ndarray, d1: np.ndarray):
    """Tie-aware float64 restatement of A.8 used to classify disagreements:
    returns sim-free nn12, nn21 (lowest index on ties) and the top-1/top-2 gap per row."""
    sim = d0.astype(np.float64) @ d1.astype(np.float64).T
    nn12 = sim.argmax(1)
    nn21 = sim.argmax(0)
    part = np.partition(sim, -2, axis=1) if sim.shape[1] > 1 else None
    gap = (part[:, -1] - part[:, -2]) if part is not None else np.full(sim.shape[0], np.inf)
    return nn12, nn21, gap, sim.max(1)


# ---------------------------------------------------------------------------------------------- input leg
def resize_target(h: int, w: int, resize_max=None, resize_force=False):
    """extract_localization.py:171-175: (h_new, w_new) of ImageDataset's resize, or (h, w) when none happens."""
    if resize_max and (resize_force or max(w, h) > resize_max):
        scale = resize_max / max(h, w)
        return int(round(h * scale)), int(round(w * scale))
    return h, w


def _cubic_coeffs(x: np.ndarray) -> np.ndarray:
    """OpenCV interpolateCubic (imgproc/resize.cpp; the reference's requirements.txt:12 pins opencv-python 4.5.5.64, this image
    has 4.13.0 - same float cubic path),
    A = -0.75, float32 arithmetic, c3 = 1 - c0 - c1 - c2."""
    x = x.astype(np.float32)
    A = np.float32(-0.75)
    one = np.float32(1)
    x1 = x + one
    c0 = ((A * x1 - np.float32(5) * A) * x1 + np.float32(8) * A) * x1 - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    y = one - x
    c2 = ((A + np.float32(2)) * y - (A + np.float32(3))) * y * y + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], -1).astype(np.float32)


def cv2_resize_cubic_f32(img: np.ndarray, w_new: int, h_new: int) -> np.ndarray:
    """cv2.resize(img_float32_HWC, (w_new, h_new), interpolation=cv2.INTER_CUBIC) restated with numpy float32
    operations in the scalar code's order: horizontal 4-tap pass on the clamped source rows, then the vertical
    4-tap combination; no rounding, no saturation.  Pinned against cv2 itself in tests/test_preprocess.py."""
    h, w = img.shape[:2]
    img = img.astype(np.float32)

    def axis(n_dst, n_src):
        scale = 1.0 / (float(n_dst) / float(n_src))
        f = ((np.arange(n_dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        frac = f - s.astype(np.float32)
        idx = np.clip(s[:, None] - 1 + np.arange(4)[None], 0, n_src - 1)
        return idx, _cubic_coeffs(frac)
    xi, xa = axis(w_new, w)
    yi, yb = axis(h_new, h)
    # horizontal pass on every source row that is needed
    hor = img[:, xi[:, 0]] * xa[None, :, 0, None]
    for k in range(1, 4):
        hor = hor + img[:, xi[:, k]] * xa[None, :, k, None]
    out = hor[yi[:, 0]] * yb[:, 0, None, None]
    for k in range(1, 4):
        out = out + hor[yi[:, k]] * yb[:, k, None, None]
    return out.astype(np.float32)


def image_dataset_item(bgr_u8: np.ndarray, resize_max=None, resize_force=False, use_cv2=False):
    """ImageDataset.__getitem__ (extract_localization.py:158-190) from the decoded BGR uint8 image on:
    -> {'image': float32 [3, h', w'] RGB / 255 (not clamped), 'original_size': [w, h]}."""
    image = bgr_u8[:, :, ::-1].astype(np.float32)
    h, w = image.shape[:2]
    hn, wn = resize_target(h, w, resize_max, resize_force)
    if (hn, wn) != (h, w):
        if use_cv2:
            import cv2
            image = cv2.resize(image, (wn, hn), interpolation=cv2.INTER_CUBIC)
        else:
            image = cv2_resize_cubic_f32(image, wn, hn)
    image = image.transpose((2, 0, 1))
    image = image / 255.
    return {"image": image.astype(np.float32), "original_size": np.array([w, h])}
