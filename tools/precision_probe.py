"""Measure tcgen05 accumulation error in isolation: inputs exactly representable in fp16, so any
difference from the float64 result is the tensor core's fp32 accumulation (and nothing else)."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import gpu_util
import torch, torch.nn.functional as F

def ref64(x, w, b, stride=1):
    y = F.conv2d(torch.from_numpy(x.transpose(2, 0, 1)[None].copy()).double(), torch.from_numpy(w).double(),
                 torch.from_numpy(b).double(), stride=stride, padding=w.shape[2] // 2)
    return y[0].permute(1, 2, 0).numpy()

rng = np.random.RandomState(0)
for (cin, k) in [(64, 1), (256, 1), (64, 3), (256, 3)]:
    H, W, cout = 24, 40, 64
    x = np.maximum(rng.randn(H, W, cin), 0).astype(np.float16).astype(np.float32) * 2
    w = (rng.randn(cout, cin, k, k) / np.sqrt(cin * k * k)).astype(np.float16).astype(np.float32)
    b = np.zeros(cout, np.float32)
    r = ref64(x, w, b)
    for prec in ["fp32", "fast", "exact"]:
        y = gpu_util.debug_conv(x, w, b, 1, 1, 0, prec)
        e = (y - r)
        print(f"K={cin*k*k:5d} {prec:5s}: mean err {e.mean():+.3e}  rms {np.sqrt((e**2).mean()):.3e}  max {np.abs(e).max():.3e}  "
              f"mean|ref| {np.abs(r).mean():.2f}  (err sign vs ref sign corr {np.mean(np.sign(e)*np.sign(r)):+.3f})", flush=True)
# exact mode with generic fp32 inputs, with and without a power-of-two weight scale
for scale in [1.0, 64.0, 4096.0]:
    cin, k, H, W, cout = 256, 3, 24, 40, 64
    x = (np.maximum(rng.randn(H, W, cin), 0) * 2).astype(np.float32)
    w = (rng.randn(cout, cin, k, k) / np.sqrt(cin * k * k)).astype(np.float32)
    b = np.zeros(cout, np.float32)
    r = ref64(x, w, b)
    y = gpu_util.debug_conv(x, w * np.float32(scale), b, 1, 1, 0, "exact") / scale
    e = y - r
    print(f"exact generic inputs, weight scale {scale:6.0f}: mean {e.mean():+.3e} rms {np.sqrt((e**2).mean()):.3e} max {np.abs(e).max():.3e}", flush=True)
    y = gpu_util.debug_conv(x, w, b, 1, 1, 0, "fp32")
    e = y - r
    print(f"fp32  generic inputs                      : mean {e.mean():+.3e} rms {np.sqrt((e**2).mean()):.3e} max {np.abs(e).max():.3e}", flush=True)
