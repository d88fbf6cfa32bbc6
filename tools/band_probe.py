import os, sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from sfd2_b200 import Extractor
from sfd2_b200.synth import synth_image_u8
REPO="/root/repo"
ex = Extractor(os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz"), precision="mixed", topk=4096)
u8 = synth_image_u8(3, 1200, 1600)
f = torch.from_numpy(np.ascontiguousarray(u8.transpose(2,0,1))[None].astype(np.float32)/255.0).pin_memory()
for i in range(5): ex.extract_host(f)
torch.cuda.synchronize()
t=time.perf_counter()
N=60
for i in range(N): ex.extract_host(f)
torch.cuda.synchronize()
dt=(time.perf_counter()-t)/N
print(os.environ.get("SFD2_HOST_BANDS","4"), os.environ.get("SFD2_BAND_LAYERS","2"), "bands,layers: extract_host f32 n=1: %.3f ms" % (dt*1e3))
sys.exit(0)
d = torch.empty_like(f, device="cuda")
for pieces in (1, 4, 12):
    fl, dl = f.view(-1), d.view(-1)
    n = fl.numel(); step = (n + pieces - 1)//pieces
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for r in range(10):
        for c in range(0, n, step): dl[c:c+step].copy_(fl[c:c+step], non_blocking=True)
    b.record(); torch.cuda.synchronize()
    print("H2D 23MB in", pieces, "pieces: %.3f ms" % (a.elapsed_time(b)/10))
