"""Which UMMA descriptor settings let a 3x3 tap read a shifted view of one TMA halo tile?"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from sfd2_b200 import _lib
torch.cuda.init(); torch.zeros(1).cuda()
lib = _lib.lib()
if not hasattr(lib, 'sfd2_debug_umma_probe'):
    sys.exit('build the library with the probes first: SFD2_WITH_PROBES=1 python -m sfd2_b200.build --force')
lib.sfd2_debug_umma_probe.restype = __import__('ctypes').c_int
for pitch in (10, 16):
    for bo in (0, 1):
        ok_all = True
        bad = []
        for ky in range(3):
            for kx in range(3):
                res = []
                for pattern in (0, 1):
                    out = np.zeros((128, 64), np.float32)
                    rc = lib.sfd2_debug_umma_probe(pitch, ky, kx, bo, pattern, out.ctypes.data_as(__import__('ctypes').c_void_p))
                    if rc != 0:
                        print("rc", rc, lib.sfd2_last_error()); sys.exit(1)
                    r = np.arange(128)
                    h, w = r // 8, r % 8
                    if pattern == 0:
                        exp = np.repeat(((h + ky) * pitch + (w + kx))[:, None], 64, 1).astype(np.float32)
                    else:
                        exp = np.repeat(np.arange(64)[None], 128, 0).astype(np.float32)
                    res.append(np.array_equal(out, exp))
                    if not res[-1] and len(bad) < 2:
                        bad.append((ky, kx, pattern, out[:10, :4].tolist()))
                ok_all &= all(res)
                print(f"pitch {pitch} base_offset {bo} tap ({ky},{kx}): rows ok={res[0]} cols ok={res[1]}")
        print(f"==> pitch {pitch} base_offset {bo}: {'ALL OK' if ok_all else 'FAIL'}", bad[:1])
