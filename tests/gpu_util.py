"""Helpers shared by the GPU parity tests (all calls go through the C ABI)."""
import ctypes as C
import os

import numpy as np

from sfd2_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WEIGHTS = os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz")

_model = {}


def model(precision="exact", use_stability=True):
    from sfd2_b200 import get_model
    key = (precision, use_stability)
    if key not in _model:
        m, _ = get_model("ressegnetv2", WEIGHTS, use_stability=use_stability, precision=precision)
        _model[key] = m.cuda()
    return _model[key]


def debug_conv(x_hwc, w_oihw, b, stride, groups, relu, precision):
    """One conv layer through sfd2_debug_conv: x [H,W,Cin] fp32 -> y [Ho,Wo,Cout] fp32."""
    ctx = model("exact").ctx
    x = np.ascontiguousarray(x_hwc, np.float32)
    w = np.ascontiguousarray(w_oihw, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    H, W, cin = x.shape
    cout, _, k, _ = w.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    y = np.zeros((Ho, Wo, cout), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(_lib.lib().sfd2_debug_conv(ctx.handle, p(x), H, W, cin, p(w), p(b), cout, k, stride, groups,
                                          int(relu), _lib.PREC[precision], p(y)), "sfd2_debug_conv")
    return y


def nms_select(heat: np.ndarray, conf_th=0.001, border=4, topk=4096, want_nms=True):
    """sfd2_nms_select_dev on a host heat-map -> (xy int [n,2], scores f32 [n], nms f32 [H,W])."""
    import torch
    ctx = model("exact").ctx
    H, W = heat.shape
    h = torch.from_numpy(np.ascontiguousarray(heat, np.float32)).cuda()
    kp = torch.zeros(topk, 2, dtype=torch.float32, device="cuda")
    sc = torch.zeros(topk, dtype=torch.float32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    nms = torch.zeros(H, W, dtype=torch.float32, device="cuda") if want_nms else None
    p = _lib.ExtractParams(conf_th=conf_th, nms_radius=4, border=border, topk=topk, precision=1, use_stability=1)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().sfd2_nms_select_dev(ctx.handle, h.data_ptr(), H, W, C.byref(p), kp.data_ptr(),
                                              sc.data_ptr(), cnt.data_ptr(), nms.data_ptr() if want_nms else None, st),
               "sfd2_nms_select_dev")
    torch.cuda.synchronize()
    n = int(cnt.item())
    return kp[:n].cpu().numpy().astype(np.int64), sc[:n].cpu().numpy(), (nms.cpu().numpy() if want_nms else None)
