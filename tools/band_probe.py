"""Single host image through sfd2_extract_host: pinned vs pageable caller memory (row-band upload, api.cu)."""
import os, sys, time, numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from sfd2_b200 import Extractor
from sfd2_b200.synth import synth_image_u8
ex = Extractor(os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz"), precision="mixed", topk=4096)
u8 = synth_image_u8(3, 1200, 1600)
f = torch.from_numpy(np.ascontiguousarray(u8.transpose(2, 0, 1))[None].astype(np.float32) / 255.0)
for name, t in (("pinned", f.pin_memory()), ("pageable", f)):
    for i in range(5): ex.extract_host(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    N = 60
    for i in range(N): ex.extract_host(t)
    torch.cuda.synchronize()
    print(os.environ.get("SFD2_HOST_BANDS", "4"), "bands", os.environ.get("SFD2_BAND_LAYERS", "2"), f"layers, {name}: extract_host f32 n=1: %.3f ms" % ((time.perf_counter() - t0) / N * 1e3))
    os.environ.pop("SFD2_DEBUG_BANDS", None)
