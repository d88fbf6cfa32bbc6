"""First-contact GPU probe: runs each check in its own subprocess under a timeout so a
hung tcgen05 kernel cannot take the whole call down, and prints numeric errors.
    python tools/gpu_probe.py            # all checks
    python tools/gpu_probe.py <check>    # one check, in-process
"""
import os
import subprocess
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def _conv_ref(x, w, b, stride, groups, relu):
    import torch
    import torch.nn.functional as F
    xt = torch.from_numpy(x.transpose(2, 0, 1)[None].copy()).double()
    y = F.conv2d(xt, torch.from_numpy(w).double(), torch.from_numpy(b).double(), stride=stride,
                 padding=w.shape[2] // 2, groups=groups)
    if relu:
        y = F.relu(y)
    return y[0].permute(1, 2, 0).numpy()


def conv_case(prec, H, W, cin, cout, k, stride, groups=1, relu=1, seed=0):
    import gpu_util
    rng = np.random.RandomState(seed)
    x = np.maximum(rng.randn(H, W, cin), 0).astype(np.float32) * 3
    w = (rng.randn(cout, cin // groups, k, k) / np.sqrt(cin // groups * k * k)).astype(np.float32)
    b = rng.randn(cout).astype(np.float32) * 0.1
    t = time.time()
    y = gpu_util.debug_conv(x, w, b, stride, groups, relu, prec)
    ref = _conv_ref(x, w, b, stride, groups, relu)
    err = float(np.abs(y - ref).max())
    scale = float(np.abs(ref).max())
    print(f"conv {prec:5s} {H}x{W} cin{cin} cout{cout} k{k} s{stride} g{groups}: max|err|={err:.3e} "
          f"(ref max {scale:.2f})  {time.time()-t:.2f}s", flush=True)
    return err, scale


CHECKS = {}


def check(fn):
    CHECKS[fn.__name__] = fn
    return fn


@check
def simt_conv():
    for args in [(40, 56, 64, 64, 3, 1), (41, 57, 64, 128, 3, 2), (24, 40, 256, 256, 1, 1), (30, 34, 256, 65, 3, 1)]:
        conv_case("fp32", *args)
    conv_case("fp32", 24, 40, 256, 256, 3, 1, groups=32)


@check
def tc_conv_s1():
    conv_case("fast", 40, 56, 64, 64, 3, 1)
    conv_case("exact", 40, 56, 64, 64, 3, 1)
    conv_case("exact", 24, 40, 128, 256, 3, 1)
    conv_case("exact", 24, 40, 256, 256, 3, 1)


@check
def tc_conv_1x1():
    conv_case("exact", 24, 40, 256, 256, 1, 1)
    conv_case("fast", 24, 40, 256, 128, 1, 1)


@check
def tc_conv_s2():
    conv_case("exact", 40, 56, 64, 64, 3, 2)
    conv_case("exact", 41, 57, 128, 128, 3, 2)
    conv_case("exact", 40, 56, 256, 256, 3, 2)


@check
def tc_conv_heads():
    conv_case("exact", 30, 34, 256, 65, 3, 1, relu=0)
    conv_case("exact", 30, 34, 256, 128, 3, 1, relu=0)


@check
def tc_conv_grouped():
    conv_case("exact", 24, 40, 256, 256, 3, 1, groups=32)
    conv_case("fast", 24, 40, 256, 256, 3, 1, groups=32)


@check
def nms():
    import gpu_util
    g = np.load(os.path.join(REPO, "tests", "golden", "nms_cases.npz"))
    for k in [f[3:] for f in g.files if f.startswith("in_")]:
        xy, sc, out = gpu_util.nms_select(g["in_" + k], conf_th=0.001, border=0, topk=8192)
        same = np.array_equal(out, g["out_" + k])
        print(f"nms {k}: bit-exact={same}  survivors={int((out > 0).sum())} selected={len(sc)}", flush=True)


def _extract(prec, name):
    import torch
    import gpu_util
    from sfd2_b200 import extract_resnet_return
    from sfd2_b200.synth import synth_image
    g = np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))
    H, W, K = int(g["H"]), int(g["W"]), int(g["K"])
    img = (g["image_u8"].astype(np.float32) / np.float32(255)).transpose(2, 0, 1)[None].copy() if "image_u8" in g.files \
        else synth_image(int(g["seed"]), H, W)
    m = gpu_util.model(prec)
    t = time.time()
    out = extract_resnet_return(m, torch.from_numpy(img), topK=K, conf_th=0.001, scales=[1.0])
    dt = time.time() - t
    kp = out["keypoints"].astype(np.int64)
    ref = g["kp_xy"].astype(np.int64)
    a = set(map(tuple, kp)); b = set(map(tuple, ref))
    common = len(a & b)
    same_order = kp.shape == ref.shape and np.array_equal(kp, ref)
    msg = f"extract {prec:5s} {name}: n={len(kp)} ref={len(ref)} common={common} same_order={same_order}"
    if same_order:
        msg += f" max|dscore|={np.abs(out['scores']-g['scores']).max():.2e} max|ddesc|={np.abs(out['descriptors']-g['desc']).max():.2e}"
    else:
        idx = {tuple(k): i for i, k in enumerate(ref)}
        pairs = [(i, idx[tuple(k)]) for i, k in enumerate(kp) if tuple(k) in idx]
        if pairs:
            i0 = np.array([p[0] for p in pairs]); i1 = np.array([p[1] for p in pairs])
            msg += f" (common) max|dscore|={np.abs(out['scores'][i0]-g['scores'][i1]).max():.2e} max|ddesc|={np.abs(out['descriptors'][i0]-g['desc'][i1]).max():.2e}"
    print(msg + f"  {dt:.2f}s", flush=True)
    if name.startswith("small") or name.startswith("odd"):
        H4, W4 = (H + 3) // 4, (W + 3) // 4
        try:
            heat = m.debug_fetch("heat", (H, W))
            print(f"   heat max|err|={np.abs(heat-g['heat']).max():.3e} (max {g['heat'].max():.3f})", flush=True)
            dm = m.debug_fetch("desc_map", (((H - 1) // 2) // 2 + 1, ((W - 1) // 2) // 2 + 1, 128))
            print(f"   desc_map max|err|={np.abs(dm.transpose(2,0,1)-g['desc_map']).max():.3e}", flush=True)
        except Exception as e:  # noqa
            print("   debug fetch failed:", e)


@check
def extract_fp32():
    for n in ["small_96x128", "odd_100x141", "c1_640x480"]:
        _extract("fp32", n)


@check
def extract_fp32_c2():
    _extract("fp32", "c2_1600x1200")


@check
def extract_exact():
    for n in ["small_96x128", "odd_100x141", "c1_640x480", "c2_1600x1200"]:
        _extract("exact", n)


@check
def extract_fast():
    for n in ["small_96x128", "c1_640x480", "c2_1600x1200"]:
        _extract("fast", n)


def _match(prec):
    import torch
    from sfd2_b200 import NearestNeighbor, Matcher, matcher_confs
    g = np.load(os.path.join(REPO, "tests", "golden", "match_cases.npz"))
    for tag in ["sq", "wide", "tall", "one", "col"]:
        d0, d1 = g[f"{tag}_d0"], g[f"{tag}_d1"]
        nn = NearestNeighbor({"do_mutual_check": True, "precision": prec})
        out = nn({"descriptors0": torch.from_numpy(d0.T.copy())[None].cuda(), "descriptors1": torch.from_numpy(d1.T.copy())[None].cuda()})
        m0 = out["matches0"][0].cpu().numpy()
        agree = (m0 == g[f"{tag}_hloc_m0"]).mean()
        ds = np.abs(out["matching_scores0"][0].cpu().numpy() - g[f"{tag}_hloc_s0"]).max()
        print(f"match {prec:5s} {tag}: hloc agree={agree:.4f} max|dscore|={ds:.2e}", flush=True)
        if f"{tag}_itloc_m0" in g.files:
            mt = Matcher(matcher_confs["NNM"], precision=prec)
            o2 = mt({"descriptors0": d0.astype(np.float64), "descriptors1": d1.astype(np.float64)})
            print(f"   itloc agree={(o2['matches0'] == g[f'{tag}_itloc_m0']).mean():.4f} "
                  f"max|dscore|={np.abs(o2['matching_scores0'] - g[f'{tag}_itloc_s0']).max():.2e}", flush=True)
    for name in ["c1_640x480", "c2_1600x1200"]:
        c = np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))
        nn = NearestNeighbor({"do_mutual_check": True, "precision": prec})
        out = nn({"descriptors0": torch.from_numpy(c["desc"].T.copy())[None].cuda(), "descriptors1": torch.from_numpy(c["desc_b"].T.copy())[None].cuda()})
        m0 = out["matches0"][0].cpu().numpy()
        print(f"match {prec:5s} {name}: agree={(m0 == c['hloc_matches0']).mean():.5f} matched={int((m0>=0).sum())} "
              f"ref={int((c['hloc_matches0']>=0).sum())}", flush=True)


@check
def match_fp32():
    _match("fp32")


@check
def match_exact():
    _match("exact")


@check
def match_fast():
    _match("fast")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            CHECKS[name]()
        sys.exit(0)
    for name in CHECKS:
        print(f"===== {name}", flush=True)
        t = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=240,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print(r.stdout[-3000:], flush=True)
            print(f"===== {name}: exit {r.returncode} in {time.time()-t:.1f}s", flush=True)
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            print(out[-3000:], flush=True)
            print(f"===== {name}: TIMEOUT", flush=True)
