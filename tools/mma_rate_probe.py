"""Cycles per SMEM-operand tcgen05.mma at M = 128 as a function of N and operand kind (sfd2_debug_mma_rate).
One CTA alone gives the architectural floor; all SMs together show what the power cap leaves of it."""
import ctypes as C
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from sfd2_b200 import _lib

torch.cuda.init(); torch.zeros(1).cuda()
lib = _lib.lib()
if not hasattr(lib, 'sfd2_debug_mma_rate'):
    sys.exit('build the library with the probes first: SFD2_WITH_PROBES=1 python -m sfd2_b200.build --force')
lib.sfd2_debug_mma_rate.restype = C.c_int
lib.sfd2_debug_mma_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
nsm = torch.cuda.get_device_properties(0).multi_processor_count
iters = 8192
for grid in (1, nsm):
    for kind, kname, kdim in ((0, "f16", 16), (1, "f8f6f4(e4m3)", 32)):
        for n in (16, 32, 64, 96, 128, 192, 256):
            out = np.zeros(grid, np.uint64)
            lib.sfd2_debug_mma_rate(n, kind, 64, grid, out.ctypes.data_as(C.c_void_p))      # warm-up
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.sfd2_debug_mma_rate(n, kind, iters, grid, out.ctypes.data_as(C.c_void_p))
            e1.record(); torch.cuda.synchronize()
            if rc != 0:
                print("rc", rc, lib.sfd2_last_error()); sys.exit(1)
            cyc = float(np.median(out)) / iters
            macs = 128 * n * kdim
            print(f"grid {grid:4d} kind {kname:13s} N {n:3d}: {cyc:7.1f} cycles/MMA  = {macs / cyc:7.0f} MAC/clk/SM"
                  f"  (floor 128*N/256 = {128 * n / 256:.0f})", flush=True)
