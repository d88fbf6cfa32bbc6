"""Drop-ins for the reference's extraction interface (same names, arguments, outputs):

  get_model(model_name, weight_path, use_stability)  <- extract_localization.py:208-218
  extract_resnet_return(model, img, conf_th, mask, topK, scales=...)  <- nets/extractor.py:97-337
  ResSegNetV2(outdim, require_stability)  <- nets/sfd2.py:259 (inference surface only)

The network, NMS / top-K and descriptor sampling all run in libsfd2_b200.so; this
module only marshals arguments and packs the float64 numpy dict the hloc / it_loc
scripts expect.  No compute happens in Python and there is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .weights import blob_from_checkpoint, fold_layers, pack_blob

__all__ = ["ResSegNetV2", "get_model", "extract_resnet_return", "Extractor"]


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class ResSegNetV2(torch.nn.Module):
    """Stands where nets.sfd2.ResSegNetV2 stands in the extraction scripts: accepts
    .eval(), .cuda(), .load_state_dict(ckpt['model'], strict=False); the forward
    pass itself lives on the device inside the native context.

    precision: 'exact' (tcgen05 fp16x3 split, parity default), 'mixed' (heat-map path
    as 'exact', descriptor head single-pass: same keypoints/scores, descriptors within
    1e-3), 'fast' (tcgen05 fp16x1) or 'fp32' (CUDA-core reference mode)."""

    def __init__(self, outdim=128, require_feature=False, require_stability=False, ms_detector=True,
                 precision="exact"):
        super().__init__()
        if outdim != 128:
            raise ValueError("only outdim=128 is supported (the shipped checkpoint)")
        if precision not in _lib.PREC:
            raise ValueError(f"precision must be one of {list(_lib.PREC)}")
        self.outdim = outdim
        self.require_stability = bool(require_stability)
        self.precision = precision
        self._blob = None
        self._ctx = None

    # -- weights -------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=False):
        sd = {k: (v.detach().cpu().double().numpy() if torch.is_tensor(v) else np.asarray(v, np.float64))
              for k, v in state_dict.items() if not k.endswith("num_batches_tracked")}
        self._blob = pack_blob(fold_layers(sd))
        self._ctx = None
        return "<All keys matched successfully>"

    def load_checkpoint(self, path):
        self._blob = blob_from_checkpoint(path)
        self._ctx = None
        return self

    # -- device --------------------------------------------------------------------------
    def cuda(self, device=None):
        if self._blob is None:
            raise _lib.Sfd2Error("load weights before .cuda()")
        if not torch.cuda.is_available():
            raise _lib.Sfd2Error("no CUDA device: sfd2_b200 has no CPU path")
        dev = _lib.device_index(device)
        if self._ctx is None or self._ctx.device != dev:
            self._ctx = _lib.Context(self._blob, dev)
        return self

    def to(self, device=None, *a, **k):
        if device is not None and torch.device(device).type == "cuda":
            return self.cuda(device)
        return self

    @property
    def ctx(self) -> "_lib.Context":
        if self._ctx is None:
            self.cuda()
        return self._ctx

    def forward(self, *a, **k):
        raise NotImplementedError("training forward is out of scope; use extract_resnet_return / Extractor")

    def prefetch(self, img):
        """Start the host -> device copy of the NEXT image now, on a side stream, so that it overlaps the extraction of the
        current one.  The next extract_resnet_return(model, img, ...) call with this same tensor finds it on the device.
        One line in the reference loop (extract_localization.py:240): call it on item i+1 before extracting item i.
        img: CPU float tensor [1,3,H,W] (pinned memory makes the copy asynchronous)."""
        t = torch.as_tensor(img)
        if t.is_cuda:
            return
        dev = torch.device("cuda", self.ctx.device)
        if self.__dict__.get("_pf_stream") is None:
            self.__dict__["_pf_stream"] = torch.cuda.Stream(dev)
        with torch.cuda.stream(self._pf_stream):
            d = t.to(dev, dtype=torch.float32, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._pf_stream)
        pf = self.__dict__.setdefault("_pf", {})
        while len(pf) >= 3:                      # a few images in flight at most (23 MB each at 1600 x 1200)
            pf.pop(next(iter(pf)))
        pf[t.data_ptr()] = (t.numel(), d, ev, t)

    def _upload_pageable(self, img):
        """Pageable host image -> device: the driver's own pageable path moves 23 MB at ~10 GB/s.  Here the image goes
        through four persistent pinned chunks: torch's multi-threaded host copy of chunk i+1 overlaps the DMA of chunk i."""
        dev = torch.device("cuda", self.ctx.device)
        flat = img.contiguous().view(-1)
        n = flat.numel()
        st = self.__dict__.get("_pg_stage")
        if st is None or st[0].numel() < n:
            st = (torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32, device=dev))
            self.__dict__["_pg_stage"] = st
        pin, out = st[0][:n], st[1][:n]
        k = 4
        step = (n + k - 1) // k
        for c in range(0, n, step):
            pin[c:c + step].copy_(flat[c:c + step])                       # host -> pinned (threaded memcpy)
            out[c:c + step].copy_(pin[c:c + step], non_blocking=True)     # pinned -> device (DMA, asynchronous)
        return out.view(img.shape)

    def _take_prefetched(self, img):
        pf = self.__dict__.get("_pf")
        if not pf or img.is_cuda:
            return None
        e = pf.get(img.data_ptr())
        if e is None or e[0] != img.numel():
            return None
        del pf[img.data_ptr()]
        d, ev = e[1], e[2]
        torch.cuda.current_stream(d.device).wait_event(ev)
        d.record_stream(torch.cuda.current_stream(d.device))
        return d.reshape(-1, *d.shape[-2:])

    def _dev_outputs(self, cap, dev):
        st = self.__dict__.get("_dev_out")
        if st is None or st[0].shape[0] < cap or st[0].device != dev:
            st = (torch.empty((cap, 2), dtype=torch.float32, device=dev), torch.empty((cap,), dtype=torch.float32, device=dev),
                  torch.empty((cap, _lib.DESC_DIM), dtype=torch.float32, device=dev), torch.empty((1,), dtype=torch.int32, device=dev))
            self.__dict__["_dev_out"] = st
        return st[0][:cap], st[1][:cap], st[2][:cap], st[3]

    def _host_staging(self, cap):
        """Pinned (kpts, scores, desc, count) buffers of the host entry point, reused across calls."""
        st = self.__dict__.get("_staging")
        if st is None or st[0].shape[0] < cap:
            st = (torch.empty((cap, 2), dtype=torch.float32).pin_memory(), torch.empty((cap,), dtype=torch.float32).pin_memory(),
                  torch.empty((cap, _lib.DESC_DIM), dtype=torch.float32).pin_memory(), torch.empty((1,), dtype=torch.int32).pin_memory())
            self.__dict__["_staging"] = st
        return st[0][:cap], st[1][:cap], st[2][:cap], st[3]

    # -- test hook -----------------------------------------------------------------------
    def debug_fetch(self, name: str, shape) -> np.ndarray:
        out = np.empty(int(np.prod(shape)), np.float32)
        n = _lib.lib().sfd2_debug_fetch(self.ctx.handle, name.encode(), _np_ptr(out), out.size)
        if n < 0:
            _lib.check(int(n), f"sfd2_debug_fetch({name})")
        return out[:n].reshape(shape)


def get_model(model_name, weight_path, use_stability=False, precision="exact"):
    """extract_localization.get_model: -> (model, extractor)."""
    if model_name != "ressegnetv2":
        raise ValueError(f"model '{model_name}' is outside the hot path (only 'ressegnetv2')")
    model = ResSegNetV2(outdim=128, require_stability=use_stability, precision=precision).eval()
    model.load_checkpoint(weight_path)
    return model, extract_resnet_return


def _params(model, conf_th, topk, border=4, border_wh=(0, 0)):
    return _lib.ExtractParams(conf_th=float(conf_th), nms_radius=4, border=border, topk=int(topk),
                              precision=_lib.PREC[model.precision], use_stability=int(model.require_stability),
                              border_w=int(border_wh[0]), border_h=int(border_wh[1]))


def _same_device(ctx, dev):
    """Foreign-device pointers would reach the kernels as illegal addresses: refuse them here."""
    if dev.type != "cuda" or dev.index != ctx.device:
        raise _lib.Sfd2Error(f"tensor lives on {dev} but the model's context is bound to cuda:{ctx.device}; "
                             "move the tensor or call model.cuda(device)")


def _pack(kp, sc, de, n):
    """-> the reference's dict: fresh, writable float64 arrays (nets/extractor.py:328-337).  The 4096 x 128 descriptor
    block is converted by torch's multi-threaded cast (numpy's scalar f32 -> f64 loop took 0.35 ms per call, more than
    the result's D2H copy); the returned array owns that memory through its base."""
    if n > 1024:
        desc = torch.from_numpy(de[:n]).to(torch.float64).numpy()
    else:
        desc = np.array(de[:n], dtype=float)
    return {"keypoints": np.array(kp[:n], dtype=float), "descriptors": desc, "scores": np.array(sc[:n], dtype=float)}


def extract_resnet_return(model, img, conf_th=0.001, mask=None, topK=-1, **kwargs):
    """nets/extractor.py:97: img float [1,3,H,W] (or [3,H,W]) RGB in [0,1], CPU or CUDA tensor.
    Returns {"keypoints": f64[K,2] (x,y), "descriptors": f64[K,128], "scores": f64[K]},
    score-descending (ties by pixel index).  0 keypoints -> empty arrays (the reference
    crashes there)."""
    if mask is not None:
        raise NotImplementedError("the reference's mask/label branch is unreachable (labels undefined, :314)")
    scales = list(kwargs.get("scales", [1.0]))
    if scales != [1.0]:
        return _extract_multiscale(model, img, conf_th, topK, scales)
    img = torch.as_tensor(img)
    if img.dtype != torch.float32:
        img = img.float()
    img = img.reshape(-1, *img.shape[-2:])
    if img.shape[0] != 3:
        raise ValueError(f"expected a 3-channel image, got {tuple(img.shape)}")
    H, W = int(img.shape[1]), int(img.shape[2])
    ctx = model.ctx
    cap = int(topK) if topK and topK > 0 else (H * W) // 16 + 4096
    p = _params(model, conf_th, cap)
    lib = _lib.lib()
    pre = model._take_prefetched(img)
    if pre is not None:
        img = pre                       # uploaded ahead of time by model.prefetch(img): no H2D on the critical path
    elif not img.is_cuda and not img.is_pinned() and img.numel() >= (1 << 20):
        img = model._upload_pageable(img)   # what a plain DataLoader yields: staged through pinned chunks, copies pipelined
    if img.is_cuda:
        img = img.contiguous()
        dev = img.device
        _same_device(ctx, dev)
        # every row of the outputs is written by the kernels (rows beyond the count read zero): no fills
        kp_d, sc_d, de_d, cnt_d = model._dev_outputs(cap, dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.sfd2_extract_dev(ctx.handle, img.data_ptr(), _lib.IMG_F32_NCHW, 1, H, W, C.byref(p),
                                        kp_d.data_ptr(), sc_d.data_ptr(), de_d.data_ptr(), cnt_d.data_ptr(), st),
                   "sfd2_extract_dev")
        kp, sc, de, cnt = model._host_staging(cap)
        kp.copy_(kp_d, non_blocking=True); sc.copy_(sc_d, non_blocking=True)
        de.copy_(de_d, non_blocking=True); cnt.copy_(cnt_d, non_blocking=True)
        _lib.check(lib.sfd2_extract_status(ctx.handle, st), "sfd2_extract_dev")     # synchronises; raises on candidate overflow
        return _pack(kp.numpy(), sc.numpy(), de.numpy(), int(cnt[0]))
    # host path: results land in persistent PINNED staging (asynchronous D2H, no per-call 2 MB allocations); _pack copies
    # them into the fresh float64 arrays the caller owns
    img = img.contiguous()
    kp, sc, de, cnt = model._host_staging(cap)
    _lib.check(lib.sfd2_extract_host(ctx.handle, img.data_ptr(), _lib.IMG_F32_NCHW, 1, H, W, C.byref(p),
                                     kp.data_ptr(), sc.data_ptr(), de.data_ptr(), cnt.data_ptr()), "sfd2_extract_host")
    return _pack(kp.numpy(), sc.numpy(), de.numpy(), int(cnt[0]))


def _extract_multiscale(model, img, conf_th, topK, scales):
    """The reference's scale loop (nets/extractor.py:118-220): for every scale the image is bilinearly
    resized (the reference resizes the normalised image; normalisation is affine per channel, so resizing the
    raw image and normalising on the device is the same map up to fp32 rounding), extracted at that size with
    the border test against the ORIGINAL extents (:181-182), keypoints are mapped back with x*W/nw, y*H/nh in
    float32 (:214-215), and the union is cut to topK by score (:322-326)."""
    import torch.nn.functional as F
    img = torch.as_tensor(img).float()
    ctx = model.ctx
    img = img.reshape(1, *img.shape[-3:]).to(torch.device("cuda", ctx.device))
    _, _, H, W = img.shape
    lib = _lib.lib()
    pts, descs, lin = [], [], []
    for si, s in enumerate(scales):
        if s == 1.0:
            cur = img
        else:
            cur = F.interpolate(img, size=(int(H * s), int(W * s)), mode="bilinear", align_corners=False)
        cur = cur.contiguous()
        nh, nw = int(cur.shape[2]), int(cur.shape[3])
        cap = int(topK) if topK and topK > 0 else (nh * nw) // 16 + 4096
        kp = torch.empty(cap, 2, dtype=torch.float32, device=img.device)
        sc = torch.empty(cap, dtype=torch.float32, device=img.device)
        de = torch.empty(cap, _lib.DESC_DIM, dtype=torch.float32, device=img.device)
        cnt = torch.empty(1, dtype=torch.int32, device=img.device)
        p = _params(model, conf_th, cap, border_wh=(W, H))
        st = torch.cuda.current_stream(img.device).cuda_stream
        _lib.check(lib.sfd2_extract_dev(ctx.handle, cur.data_ptr(), _lib.IMG_F32_NCHW, 1, nh, nw, C.byref(p),
                                        kp.data_ptr(), sc.data_ptr(), de.data_ptr(), cnt.data_ptr(), st),
                   "sfd2_extract_dev")
        _lib.check(lib.sfd2_extract_status(ctx.handle, st), "sfd2_extract_dev")
        n = int(cnt.item())
        k = kp[:n].cpu().numpy()
        lin.append((si << 40) + k[:, 1].astype(np.int64) * nw + k[:, 0].astype(np.int64))
        k = np.stack([k[:, 0] * np.float32(W) / np.float32(nw), k[:, 1] * np.float32(H) / np.float32(nh)], 1)
        pts.append(np.concatenate([k.astype(np.float32), sc[:n].cpu().numpy()[:, None]], 1))
        descs.append(de[:n].cpu().numpy())
    pts, descs, lin = np.vstack(pts), np.vstack(descs), np.concatenate(lin)
    order = np.lexsort((lin, -pts[:, 2].astype(np.float64)))
    if topK and topK > 0:
        order = order[:topK]
    return {"keypoints": np.array(pts[order, :2], dtype=float), "descriptors": np.array(descs[order], dtype=float),
            "scores": np.array(pts[order, 2], dtype=float)}


class Extractor:
    """Batched, device-resident front end for throughput work (bench / dataset sweeps):
    images stay in HBM, outputs are fixed-capacity torch tensors plus counts."""

    def __init__(self, weight_path, use_stability=True, precision="exact", topk=4096, conf_th=0.001, device=None):
        self.model = ResSegNetV2(require_stability=use_stability, precision=precision).load_checkpoint(weight_path)
        self.model.cuda(device)
        self.topk, self.conf_th = int(topk), float(conf_th)

    def __call__(self, images: torch.Tensor):
        """images: CUDA float32 [n,3,H,W] in [0,1] or CUDA uint8 [n,H,W,3]."""
        if not images.is_cuda:
            raise ValueError("Extractor takes device-resident batches; use extract_resnet_return for host images")
        images = images.contiguous()
        if images.dtype == torch.uint8:
            n, H, W, _ = images.shape
            dt = _lib.IMG_U8_NHWC
        else:
            n, _, H, W = images.shape
            dt = _lib.IMG_F32_NCHW
        dev = images.device
        _same_device(self.model.ctx, dev)
        kp = torch.empty(n, self.topk, 2, dtype=torch.float32, device=dev)
        sc = torch.empty(n, self.topk, dtype=torch.float32, device=dev)
        de = torch.empty(n, self.topk, _lib.DESC_DIM, dtype=torch.float32, device=dev)
        cnt = torch.empty(n, dtype=torch.int32, device=dev)
        p = _params(self.model, self.conf_th, self.topk)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().sfd2_extract_dev(self.model.ctx.handle, images.data_ptr(), dt, n, H, W, C.byref(p),
                                               kp.data_ptr(), sc.data_ptr(), de.data_ptr(), cnt.data_ptr(), st),
                   "sfd2_extract_dev")
        return {"keypoints": kp, "scores": sc, "descriptors": de, "counts": cnt}

    def check_status(self):
        """The batched device call is asynchronous; this synchronises the current stream and raises Sfd2Error if any
        extract since the last check produced more NMS candidates than the workspace holds (truncated results)."""
        ctx = self.model.ctx
        st = torch.cuda.current_stream(torch.device("cuda", ctx.device)).cuda_stream
        _lib.check(_lib.lib().sfd2_extract_status(ctx.handle, st), "sfd2_extract_dev")

    def extract_host(self, images):
        """Batched host entry point (sfd2_extract_host): images = CPU tensor / ndarray float32 [n,3,H,W] in
        [0,1] (pinned memory makes the copies asynchronous) or uint8 [n,H,W,3].  The H2D copy of image
        i+1 overlaps the kernels of image i; returns numpy arrays with fixed capacity + counts."""
        t = torch.as_tensor(images)
        if t.is_cuda:
            raise ValueError("extract_host takes host images; call the Extractor for device-resident batches")
        t = t.contiguous()
        if t.dtype == torch.uint8:
            n, H, W, _ = t.shape
            dt = _lib.IMG_U8_NHWC
        else:
            t = t.float()
            n, _, H, W = t.shape
            dt = _lib.IMG_F32_NCHW
        kp = self._pinned("kp", (n, self.topk, 2), torch.float32)
        sc = self._pinned("sc", (n, self.topk), torch.float32)
        de = self._pinned("de", (n, self.topk, _lib.DESC_DIM), torch.float32)
        cnt = self._pinned("cnt", (n,), torch.int32)
        p = _params(self.model, self.conf_th, self.topk)
        _lib.check(_lib.lib().sfd2_extract_host(self.model.ctx.handle, t.data_ptr(), dt, n, H, W, C.byref(p),
                                                kp.data_ptr(), sc.data_ptr(), de.data_ptr(), cnt.data_ptr()),
                   "sfd2_extract_host")
        return {"keypoints": kp.numpy(), "scores": sc.numpy(), "descriptors": de.numpy(), "counts": cnt.numpy()}

    def _pinned(self, name, shape, dtype):
        """Reusable pinned output buffers (results are overwritten by the next extract_host call)."""
        cache = self.__dict__.setdefault("_pin", {})
        t = cache.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=dtype).pin_memory()
            cache[name] = t
        return t
