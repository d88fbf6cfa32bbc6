"""Input leg end to end: JPEG files on disk -> DeviceImageLoader (decode threads, pinned ring, device preprocess) -> extractor.
    python tools/loader_bench.py [n_files] [workers]
Prints images/s of (a) the loader alone, (b) loader + batched device extraction (8 images per native call), (c) loader + the
reference-signature call per image, (d) the reference's own arrangement restated: cv2.imread + float + cv2.resize + /255 on
the main thread feeding the same extractor call (what ImageDataset does per item, extract_localization.py:158-190)."""
import os, sys, tempfile, time
import numpy as np, torch, cv2
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from sfd2_b200 import Extractor, extract_resnet_return
from sfd2_b200.preprocess import DeviceImageLoader
from sfd2_b200.synth import synth_image_u8

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
W = os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz")
ex = Extractor(W, precision="mixed", topk=4096)
tmp = tempfile.mkdtemp()
base = [synth_image_u8(s, 1500, 2000, sigma=2.0) for s in range(4)]      # 2000x1500 photos -> resize_max 1600 -> 1600x1200
names = []
for i in range(n):
    nm = f"img_{i:04d}.jpg"
    cv2.imwrite(os.path.join(tmp, nm), np.roll(base[i % 4], (7 * i, 13 * i), axis=(0, 1))[:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, 92])
    names.append(nm)
lst = os.path.join(tmp, "list.txt")
open(lst, "w").write("\n".join(names) + "\n")
conf = {"resize_max": 1600, "grayscale": False}
print(f"{n} JPEG files of 2000x1500 ({os.path.getsize(os.path.join(tmp, names[0])) / 1e6:.2f} MB each), {workers} decode threads, host threads {os.cpu_count()}")

def run(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); k = fn(); torch.cuda.synchronize()
    return k / (time.perf_counter() - t0)

def loader_only():
    k = 0
    for d in DeviceImageLoader(tmp, conf, ex.model, image_list=lst, workers=workers, depth=4):
        k += 1
    return k
def loader_batched():
    k, buf = 0, []
    for d in DeviceImageLoader(tmp, conf, ex.model, image_list=lst, workers=workers, depth=8):
        buf.append(d["image"]); k += 1
        if len(buf) == 8:
            ex(torch.cat(buf)); buf = []
    if buf:
        ex(torch.cat(buf))
    ex.check_status()
    return k
def loader_single():
    k = 0
    for d in DeviceImageLoader(tmp, conf, ex.model, image_list=lst, workers=workers, depth=4):
        extract_resnet_return(ex.model, d["image"], topK=4096, conf_th=0.001, scales=[1.0]); k += 1
    return k
def reference_arrangement():
    k = 0
    for nm in names[:max(8, n // 4)]:
        im = cv2.imread(os.path.join(tmp, nm), cv2.IMREAD_COLOR)[:, :, ::-1].astype(np.float32)
        im = cv2.resize(im, (1600, 1200), interpolation=cv2.INTER_CUBIC).transpose(2, 0, 1) / 255.
        extract_resnet_return(ex.model, torch.from_numpy(np.ascontiguousarray(im[None], dtype=np.float32)), topK=4096, conf_th=0.001, scales=[1.0]); k += 1
    return k
loader_only()
print(f"loader alone (decode + upload + device preprocess): {run(loader_only):.1f} images/s")
print(f"loader + batched device extraction (8 per call)    : {run(loader_batched):.1f} images/s")
print(f"loader + reference-signature call per image        : {run(loader_single):.1f} images/s")
print(f"main-thread cv2 decode/resize + the same call       : {run(reference_arrangement):.1f} images/s")
