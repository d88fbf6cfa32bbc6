"""Latency of the reference-signature call extract_resnet_return(model, host image) at 1600x1200 / top-4096, and of the
device-resident batch, for A/B runs (SFD2_TC_PDL=0|1 ...).   python tools/single_call.py [precision]"""
import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from sfd2_b200 import Extractor, extract_resnet_return
from sfd2_b200.synth import synth_image_u8
prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
W = os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz")
ex = Extractor(W, precision=prec, topk=4096)
u8 = np.stack([synth_image_u8(s, 1200, 1600) for s in range(4)])
f32 = [(u.astype(np.float32) / np.float32(255)).transpose(2, 0, 1)[None].copy() for u in u8]
pinned = [torch.from_numpy(x).pin_memory() for x in f32]
pageable = [torch.from_numpy(x) for x in f32]
dev = torch.from_numpy(np.concatenate(f32)).cuda()
def rate(fn, n):
    for i in range(3): fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return n / (time.perf_counter() - t0)
print("env", {k: v for k, v in os.environ.items() if k.startswith("SFD2_")})
print(f"single call, pinned host image : {rate(lambda i: extract_resnet_return(ex.model, pinned[i % 4], topK=4096, conf_th=0.001, scales=[1.0]), 40):.1f} img/s")
def pf(i):
    ex.model.prefetch(pinned[(i + 1) % 4])
    return extract_resnet_return(ex.model, pinned[i % 4], topK=4096, conf_th=0.001, scales=[1.0])
ex.model.prefetch(pinned[0])
print(f"single call + prefetch of next : {rate(pf, 40):.1f} img/s")
print(f"single call, pageable host image: {rate(lambda i: extract_resnet_return(ex.model, pageable[i % 4], topK=4096, conf_th=0.001, scales=[1.0]), 40):.1f} img/s")
print(f"single call, device image      : {rate(lambda i: extract_resnet_return(ex.model, dev[i % 4:i % 4 + 1], topK=4096, conf_th=0.001, scales=[1.0]), 40):.1f} img/s")
print(f"device batch of 1 (async)      : {rate(lambda i: ex(dev[i % 4:i % 4 + 1]), 40):.1f} img/s")
print(f"device batch of 4 (async)      : {4 * rate(lambda i: ex(dev), 20):.1f} img/s")
ctx = ex.model.ctx
ctx.profile(True); ctx.profile_read()
for i in range(4): ex(dev[i:i + 1])
pr = ctx.profile_read(); ctx.profile(False)
tot = sum(v[1] for v in pr.values()) / 4
print(f"sum of kernel times per image (serialised, events): {tot:.3f} ms")
