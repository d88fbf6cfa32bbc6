#!/usr/bin/env python
"""A/B timing of library toggles in one process (one pool of synthetic images, one context per configuration).

    python tools/ab_bench.py "name:prec:ENV=V,ENV2=V" ...

Each configuration: 3 warm-up steps, 8 timed steps of 8 device-resident 1600x1200 images (CUDA events), then 2
profiled steps for the per-layer table.  The toggles are read by sfd2_create, so a fresh context sees them.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    from sfd2_b200 import Extractor
    from sfd2_b200.synth import synth_image_u8
    H, W, B = 1200, 1600, 8
    dev = torch.device("cuda", 0)
    pool_u8 = np.stack([synth_image_u8(s, H, W) for s in range(16)])
    pool = torch.from_numpy(pool_u8).to(dev).float().div_(255.0).permute(0, 3, 1, 2).contiguous()
    weights = os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz")
    for spec in sys.argv[1:]:
        name, prec, envs = (spec.split(":") + ["", ""])[:3]
        saved = {}
        for kv in filter(None, envs.split(",")):
            k, v = kv.split("=")
            saved[k] = os.environ.get(k)
            os.environ[k] = v
        ex = Extractor(weights, use_stability=True, precision=prec, topk=4096, conf_th=0.001, device=dev)
        ctx = ex.model.ctx
        step = lambda i: ex(pool[(i % 2) * B:(i % 2) * B + B])
        for i in range(3):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(8):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (8 * B)
        ctx.profile(True); ctx.profile_read()
        for i in range(2):
            step(i)
        prof = ctx.profile_read(); ctx.profile(False)
        per = {k.split(":")[-1]: round(v[1] / (2 * B) * 1000) for k, v in prof.items()}
        print(f"## {name} [{prec}] {envs}: {ms:.3f} ms/image = {1000 / ms:.1f} images/s; sum of kernels {sum(per.values())} us")
        print("   ", per, flush=True)
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        del ex, ctx
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
