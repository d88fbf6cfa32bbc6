"""Export the reference checkpoint's tensors to an .npz the repo can carry.

Run in the build container (where /root/reference exists):
    python oracle/export_weights.py [src.pth] [dst.npz]
The npz keeps the reference's own state-dict key names and fp32 values
(`num_batches_tracked` counters dropped); nothing is folded or re-ordered here —
folding happens in sfd2_b200/weights.py at load time, exactly as it would for a
user-supplied .pth.
"""
import sys
import numpy as np
import torch

SRC = "/root/reference/weights/20220810_ressegnetv2_wapv2_ce_sd2mfsf_uspg.pth"
DST = "weights/ressegnetv2_wapv2.npz"


def main(src=SRC, dst=DST):
    ck = torch.load(src, map_location="cpu", weights_only=False)
    sd = ck["model"]
    out = {k: v.numpy() for k, v in sd.items() if not k.endswith("num_batches_tracked")}
    np.savez(dst, **out)
    n = sum(v.size for v in out.values())
    print(f"wrote {dst}: {len(out)} tensors, {n} values")


if __name__ == "__main__":
    main(*sys.argv[1:])
