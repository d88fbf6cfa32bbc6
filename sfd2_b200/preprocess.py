"""Input leg of the extraction sweep - the reference's ImageDataset (extract_localization.py:120-190) with the
arithmetic moved to the device:

  reference: 4 DataLoader worker processes each do cv2.imread -> BGR->RGB -> float32 -> cv2.resize(INTER_CUBIC) ->
             CHW -> / 255 and ship a 23 MB float tensor to the main process, which uploads it;
  here:      worker THREADS only decode (cv2.imread releases the GIL); the uint8 image (3 bytes per pixel) goes through a
             pinned staging ring to the device, where one kernel (csrc/preprocess.cu) does colour order, cubic resize,
             layout and scaling.  Decode and upload of the next images overlap the extraction of the current one.

DeviceImageLoader yields what DataLoader(ImageDataset(...)) yields - {'name': [str], 'image': float32 [1,3,h,w],
'original_size': int tensor [1,2] (w, h)} - except that 'image' already lives on the GPU, so the reference's loop body
(extract_localization.py:240-272) runs unchanged on top of it.
"""
import ctypes as C
import os
from collections import deque
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib

__all__ = ["resize_target", "preprocess_dev", "DeviceImageLoader"]


def resize_target(h, w, resize_max=None, resize_force=False):
    """extract_localization.py:171-175 -> (h_new, w_new)."""
    if resize_max and (resize_force or max(w, h) > resize_max):
        scale = resize_max / max(h, w)
        return int(round(h * scale)), int(round(w * scale))
    return h, w


def preprocess_dev(img_u8: torch.Tensor, ctx, resize_max=None, resize_force=False, bgr=True, out=None):
    """img_u8: CUDA uint8 [h,w,3] as cv2.imread returns it (BGR; bgr=False for RGB input).
    -> CUDA float32 [1,3,h',w'] = what ImageDataset.__getitem__ would have produced (RGB, cubic-resized, / 255)."""
    if not (img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.dim() == 3 and img_u8.shape[2] == 3):
        raise ValueError("preprocess_dev takes a CUDA uint8 [h,w,3] image")
    img_u8 = img_u8.contiguous()
    h, w = int(img_u8.shape[0]), int(img_u8.shape[1])
    hn, wn = resize_target(h, w, resize_max, resize_force)
    if out is None:
        out = torch.empty((1, 3, hn, wn), dtype=torch.float32, device=img_u8.device)
    st = torch.cuda.current_stream(img_u8.device).cuda_stream
    _lib.check(_lib.lib().sfd2_preprocess_dev(ctx.handle, img_u8.data_ptr(), h, w, int(bool(bgr)), hn, wn, out.data_ptr(), st),
               "sfd2_preprocess_dev")
    return out


class DeviceImageLoader:
    """for data in DeviceImageLoader(image_dir, conf['preprocessing'], model, image_list=...):  # the reference's loop body

    root / conf / image_list as ImageDataset (globs, grayscale (unsupported: outside the presets), resize_max,
    resize_force).  `workers` decode threads, `depth` images in flight."""
    default_conf = {"globs": ["*.jpg", "*.png", "*.jpeg", "*.JPG", "*.PNG"], "grayscale": False, "resize_max": None,
                    "resize_force": False}

    def __init__(self, root, conf, model, image_list=None, workers=4, depth=4):
        self.conf = conf = SimpleNamespace(**{**self.default_conf, **conf})
        if conf.grayscale:
            raise NotImplementedError("grayscale input is outside the hot path (every SFD2 preset is RGB)")
        self.root = Path(root)
        self.paths = []
        if image_list is None:
            for g in conf.globs:
                self.paths += list(self.root.glob("**/" + g))
            if len(self.paths) == 0:
                raise ValueError(f"Could not find any image in root: {root}.")
            self.paths = [i.relative_to(self.root) for i in self.paths]
        else:
            with open(image_list, "r") as f:
                self.paths = [Path(l.strip()) for l in f.readlines() if l.strip()]
        self.ctx = model.ctx
        self.device = torch.device("cuda", self.ctx.device)
        self.workers, self.depth = int(workers), max(1, int(depth))
        self._copy_stream = torch.cuda.Stream(self.device)
        self._pinned = {}
        self._slot_ev = {}

    def __len__(self):
        return len(self.paths)

    def _decode(self, path):
        import cv2
        image = cv2.imread(str(self.root / path), cv2.IMREAD_COLOR)       # BGR uint8 [h,w,3]
        if image is None:
            raise ValueError(f"Cannot read image {str(path)}.")
        return image

    def _staging(self, slot, shape):
        t = self._pinned.get(slot)
        n = int(np.prod(shape))
        if t is None or t.numel() < n:
            t = torch.empty(n, dtype=torch.uint8).pin_memory()
            self._pinned[slot] = t
        return t[:n].view(*shape)

    def _upload(self, slot, image):
        """pinned staging -> device on the copy stream, then the preprocess kernel; returns (tensor, event)."""
        prev = self._slot_ev.get(slot)
        if prev is not None:
            prev.synchronize()              # the H2D copy that last read this pinned slot has finished
        stage = self._staging(slot, image.shape)
        stage.copy_(torch.from_numpy(image))
        with torch.cuda.stream(self._copy_stream):
            dev = stage.to(self.device, non_blocking=True)
            out = preprocess_dev(dev, self.ctx, self.conf.resize_max, self.conf.resize_force, bgr=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._slot_ev[slot] = ev
        return out, ev, dev

    def __iter__(self):
        pool = ThreadPoolExecutor(max_workers=self.workers)
        try:
            pending = deque()       # decode futures, in order
            ready = deque()         # (path, size, device tensor, event, keepalive)
            it = iter(self.paths)
            slot = 0

            def refill():
                while len(pending) + len(ready) < self.depth + self.workers:
                    p = next(it, None)
                    if p is None:
                        return
                    pending.append((p, pool.submit(self._decode, p)))
            refill()
            while pending or ready:
                # move decoded images to the device as soon as they are available (keeps `depth` uploads ahead)
                while pending and len(ready) < self.depth and (pending[0][1].done() or not ready):
                    p, fut = pending.popleft()
                    image = fut.result()
                    h, w = image.shape[:2]
                    out, ev, keep = self._upload(slot % (self.depth + 1), image)
                    slot += 1
                    ready.append((p, (w, h), out, ev, keep))
                    refill()
                p, size, out, ev, keep = ready.popleft()
                torch.cuda.current_stream(self.device).wait_event(ev)
                yield {"name": [str(p)], "image": out, "original_size": torch.tensor([list(size)])}
                refill()
        finally:
            pool.shutdown(wait=False, cancel_futures=True)
