"""Deterministic synthetic inputs for the SFD2 hot path (bench + parity tests).

The survey's generator low-passes uniform noise with
cv2.GaussianBlur(sigma=3).  cv2's result depends on the host's SIMD dispatch, so
the same recipe is restated here with numpy element-wise float32 operations in a
fixed order (every add / multiply individually rounded, no FMA contraction):
the uint8 image is bit-identical on any host.  Plain uniform noise gives only a
few dozen detections above conf_th; the sigma=3 low-pass gives >K candidates at
both benchmark sizes, so top-K really truncates.
"""
import numpy as np

__all__ = ["synth_image_u8", "synth_image", "shifted_twin", "synth_descriptors"]


def _gauss_taps(sigma: float) -> np.ndarray:
    r = int(round(4.0 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return (k / k.sum()).astype(np.float32)


def _blur_axis(a: np.ndarray, taps: np.ndarray, axis: int) -> np.ndarray:
    r = (len(taps) - 1) // 2
    pad = [(0, 0)] * a.ndim
    pad[axis] = (r, r)
    p = np.pad(a, pad, mode="reflect")
    n = a.shape[axis]
    out = np.zeros_like(a)
    for t in range(len(taps)):              # fixed accumulation order
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(t, t + n)
        out = out + p[tuple(sl)] * taps[t]  # float32 mul, float32 add
    return out


def synth_image_u8(seed: int, H: int, W: int, sigma: float = 3.0) -> np.ndarray:
    """uint8 RGB image [H, W, 3]: low-passed uniform noise, min-max stretched."""
    rng = np.random.RandomState(seed)
    im = rng.rand(H, W, 3).astype(np.float32)
    taps = _gauss_taps(sigma)
    im = _blur_axis(_blur_axis(im, taps, 0), taps, 1)
    lo, hi = im.min(), im.max()
    im = (im - lo) / (hi - lo)
    return (im * np.float32(255.0)).astype(np.uint8)


def synth_image(seed: int, H: int, W: int, sigma: float = 3.0) -> np.ndarray:
    """Model input: float32 [1, 3, H, W] in [0, 1] (what ImageDataset yields,
    extract_localization.py:158-190)."""
    u8 = synth_image_u8(seed, H, W, sigma)
    return (u8.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)[None].copy()


def shifted_twin(u8: np.ndarray, dy: int = 6, dx: int = 10) -> np.ndarray:
    """Second frame of a synthetic pair: the same image rolled by (dy, dx)."""
    return np.roll(u8, (dy, dx), axis=(0, 1))


def synth_descriptors(seed: int, n: int, m: int, d: int = 128, noise: float = 0.3):
    """Matcher stress set: unit-norm gaussian rows; the first min(n, m)//2 rows of
    d1 are noisy copies of (a permutation of) d0 rows, the rest are unrelated."""
    rng = np.random.RandomState(seed)
    d0 = rng.randn(n, d).astype(np.float32)
    d1 = rng.randn(m, d).astype(np.float32)
    k = min(n, m) // 2
    perm = rng.permutation(n)[:k]
    d1[:k] = d0[perm] + noise * rng.randn(k, d).astype(np.float32)
    d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    return d0, d1
