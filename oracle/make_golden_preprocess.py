"""Golden vectors for the input leg: what cv2 itself (the reference's dependency, extract_localization.py:158-190) produces.
    python oracle/make_golden_preprocess.py        -> tests/golden/preprocess_cases.npz
Run in the build container (opencv-python 4.13); the fixture is committed, the GPU box only reads it."""
import os
import sys

import cv2
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from sfd2_b200.synth import synth_image_u8  # noqa: E402


def item(bgr, resize_max, resize_force):
    """ImageDataset.__getitem__ verbatim in behaviour (from the decoded image on), resize by cv2."""
    image = bgr[:, :, ::-1].astype(np.float32)
    size = image.shape[:2][::-1]
    w, h = size
    if resize_max and (resize_force or max(w, h) > resize_max):
        scale = resize_max / max(h, w)
        h_new, w_new = int(round(h * scale)), int(round(w * scale))
        image = cv2.resize(image, (w_new, h_new), interpolation=cv2.INTER_CUBIC)
    image = image.transpose((2, 0, 1)) / 255.
    return image.astype(np.float32), np.array(size)


out = {"cv2_version": cv2.__version__}
for tag, (h, w, rmax, force) in {"up": (60, 84, 200, True), "down": (213, 320, 128, False), "same": (48, 64, 128, False)}.items():
    bgr = np.ascontiguousarray(synth_image_u8({"up": 1, "down": 2, "same": 3}[tag], h, w, sigma=1.5)[:, :, ::-1])
    img, size = item(bgr, rmax, force)
    out.update({f"{tag}_bgr": bgr, f"{tag}_image": img, f"{tag}_original_size": size, f"{tag}_resize_max": rmax, f"{tag}_force": force})
np.savez_compressed(os.path.join(REPO, "tests", "golden", "preprocess_cases.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
