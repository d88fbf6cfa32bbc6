"""Drop-ins for the reference's two mutual-nearest-neighbour matchers:

  NearestNeighbor  <- hloc/matchers/nearest_neighbor.py:27-57 (BaseModel plugin, torch in/out)
  Matcher          <- it_loc/matcher.py:85-119 (numpy in/out, mode 'nnm')
  BaseModel        <- hloc/utils/base_model.py:7-37 (so the plugin also works standalone)

Both run the same native kernel (csrc/match.cu, csrc/tc_match.cu): similarity GEMM
with the row / column arg-max fused into its epilogue, then the mutual check.
"""
import ctypes as C
from copy import copy

import numpy as np
import torch

from . import _lib

__all__ = ["BaseModel", "NearestNeighbor", "NearestNeighborMixin", "Matcher", "confs", "match_batched", "match_one_to_many"]

_CTX = {}


def _ctx(device_index: int) -> "_lib.Context":
    """The matcher needs no network weights: one matcher-only native context per device (sfd2_create(NULL))."""
    if device_index not in _CTX:
        _CTX[device_index] = _lib.Context(None, device_index)
    return _CTX[device_index]


def _mparams(mutual, dist_th, ratio_th, precision, ratio_mode=0):
    return _lib.MatchParams(do_mutual_check=int(bool(mutual)),
                            distance_threshold=float(dist_th) if dist_th else 0.0,
                            ratio_threshold=float(ratio_th) if ratio_th else 0.0,
                            precision=_lib.PREC[precision], ratio_mode=int(ratio_mode))


def match_dev(d0: torch.Tensor, d1: torch.Tensor, mutual=True, dist_th=None, ratio_th=None, precision="exact",
              ratio_mode=0, raw=False):
    """d0 [N,D], d1 [M,D] CUDA float32 row-major -> (matches0 int32 [N], sim0 float32 [N]) on the device.
    matches0 < 0 means no match; with raw=True the native codes are kept (-1 = the row failed its own
    ratio/distance test, -2 = it failed only the mutual check)."""
    assert d0.is_cuda and d1.is_cuda and d0.dtype == torch.float32 and d1.dtype == torch.float32
    d0, d1 = d0.contiguous(), d1.contiguous()
    n0, d = d0.shape
    n1 = d1.shape[0]
    dev = d0.device
    # match_finish_kernel writes every one of the n0 rows (also when d1 is empty), so no fill kernels are needed
    m0 = torch.empty((n0,), dtype=torch.int32, device=dev)
    s0 = torch.empty((n0,), dtype=torch.float32, device=dev)
    if n0 == 0:
        return m0, s0
    # raw=False: the finish kernel itself reports every unmatched row as -1 (SFD2_MATCH_PLAIN_CODES), no clamp launch
    p = _mparams(mutual, dist_th, ratio_th, precision, int(ratio_mode) | (0 if raw else 0x100))
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().sfd2_match_dev(_ctx(dev.index or 0).handle, d0.data_ptr(), n0, d1.data_ptr(), n1, d,
                                         C.byref(p), m0.data_ptr(), s0.data_ptr(), st), "sfd2_match_dev")
    return m0, s0


def match_batched(d0: torch.Tensor, off0, d1: torch.Tensor, off1, mutual=True, dist_th=None, precision="exact"):
    """Many pairs in one native call; off0/off1 are python/numpy int sequences of length npairs+1."""
    d0, d1 = d0.contiguous(), d1.contiguous()
    o0 = np.ascontiguousarray(off0, np.int32)
    o1 = np.ascontiguousarray(off1, np.int32)
    npairs = len(o0) - 1
    dev = d0.device
    m0 = torch.full((d0.shape[0],), -1, dtype=torch.int32, device=dev)
    s0 = torch.zeros((d0.shape[0],), dtype=torch.float32, device=dev)
    p = _mparams(mutual, dist_th, None, precision)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().sfd2_match_batched_dev(_ctx(dev.index or 0).handle, d0.data_ptr(),
                                                 o0.ctypes.data_as(C.c_void_p), d1.data_ptr(),
                                                 o1.ctypes.data_as(C.c_void_p), npairs, d0.shape[1], C.byref(p),
                                                 m0.data_ptr(), s0.data_ptr(), st), "sfd2_match_batched_dev")
    m0.clamp_(min=-1)
    return m0, s0


def match_one_to_many(q: torch.Tensor, db: torch.Tensor, db_off, mutual=True, dist_th=None, precision="exact"):
    """One query set q [N,128] against ndb db sets stored back to back in db (row offsets db_off, length
    ndb+1) in ONE grouped launch - the localizer's pattern (a query against its retrieved db images).
    -> (matches0 int32 [ndb, N] local indices or -1, sim0 float32 [ndb, N])."""
    q, db = q.contiguous(), db.contiguous()
    off = np.ascontiguousarray(db_off, np.int32)
    ndb = len(off) - 1
    dev = q.device
    m0 = torch.full((ndb, q.shape[0]), -1, dtype=torch.int32, device=dev)
    s0 = torch.zeros((ndb, q.shape[0]), dtype=torch.float32, device=dev)
    p = _mparams(mutual, dist_th, None, precision)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().sfd2_match_one_to_many_dev(_ctx(dev.index or 0).handle, q.data_ptr(), q.shape[0], db.data_ptr(),
                                                     off.ctypes.data_as(C.c_void_p), ndb, q.shape[1], C.byref(p),
                                                     m0.data_ptr(), s0.data_ptr(), st), "sfd2_match_one_to_many_dev")
    m0.clamp_(min=-1)
    return m0, s0


class BaseModel(torch.nn.Module):
    """hloc/utils/base_model.py:7-37 (same contract: default_conf merge, _init, forward -> _forward)."""
    default_conf = {}
    required_data_keys = []

    def __init__(self, conf):
        super().__init__()
        self.conf = conf = {**self.default_conf, **conf}
        self.required_data_keys = copy(self.required_data_keys)
        self._init(conf)

    def forward(self, data):
        for key in self.required_data_keys:
            assert key in data, "Missing key {} in data".format(key)
        return self._forward(data)

    def _init(self, conf):
        raise NotImplementedError

    def _forward(self, data):
        raise NotImplementedError


class NearestNeighborMixin:
    """The plugin body, independent of which BaseModel it is mixed with - so a reference
    checkout can define `class NearestNeighbor(NearestNeighborMixin, BaseModel)` inside
    hloc/matchers/<name>.py and dynamic_load (base_model.py:40-50) will pick it up."""
    default_conf = {
        "ratio_threshold": None,
        "distance_threshold": None,
        "do_mutual_check": True,
        "precision": "exact",
    }
    required_inputs = ["descriptors0", "descriptors1"]

    def _init(self, conf):
        pass

    def _forward(self, data):
        d0, d1 = data["descriptors0"], data["descriptors1"]      # [B, D, N], [B, D, M]
        if not d0.is_cuda:
            raise _lib.Sfd2Error("NearestNeighbor: descriptors must be CUDA tensors (no CPU path)")
        B = d0.shape[0]
        ms, ss = [], []
        for b in range(B):
            a = d0[b].float().t().contiguous()
            c = d1[b].float().t().contiguous()
            m0, s0 = match_dev(a, c, self.conf["do_mutual_check"], self.conf["distance_threshold"],
                               self.conf["ratio_threshold"], self.conf.get("precision", "exact"), ratio_mode=0, raw=True)
            # find_nn (nearest_neighbor.py:14-15): rows failing their own ratio / distance test get score 0;
            # rows rejected only by the mutual check keep (sim + 1) / 2
            scores = torch.where(m0 == -1, s0.new_tensor(0), (s0 + 1) / 2)
            m0 = m0.clamp(min=-1)
            ms.append(m0.long())
            ss.append(scores)
        return {"matches0": torch.stack(ms), "matching_scores0": torch.stack(ss)}


class NearestNeighbor(NearestNeighborMixin, BaseModel):
    pass


confs = {   # it_loc/matcher.py:24-82, the entries on the hot path
    "NNM": {"output": "NNM", "model": {"name": "nnm", "do_mutual_check": True, "distance_threshold": None}},
    "ONN": {"output": "ONN", "model": {"name": "nn", "do_mutual_check": False, "distance_threshold": None}},
    "NNR": {"output": "NNR", "model": {"name": "nnr", "do_mutual_check": True, "distance_threshold": 0.9}},
}


class Matcher(torch.nn.Module):
    """it_loc/matcher.py:85-119: numpy [N,D] / [M,D] in (any float dtype), numpy out;
    matching_scores0 is the RAW max cosine of every row.  Mode 'nnm' only (hot path)."""

    def __init__(self, conf, precision="exact"):
        super().__init__()
        self.conf = conf
        self.mode = conf["model"]["name"]
        self.precision = precision
        if self.mode not in ("nnm", "nn", "nnr"):
            raise NotImplementedError(f"matcher mode '{self.mode}' is outside the hot path")

    def cuda(self, device=None):
        return self

    def forward(self, data):
        d0 = np.ascontiguousarray(data["descriptors0"], dtype=np.float32)
        d1 = np.ascontiguousarray(data["descriptors1"], dtype=np.float32)
        n0 = d0.shape[0]
        n1 = d1.shape[0]
        d = d0.shape[1] if d0.ndim == 2 else _lib.DESC_DIM
        m0 = np.full((n0,), -1, np.int32)
        s0 = np.zeros((n0,), np.float32)
        if n0 > 0:
            # 'nnr' = mutual NN + symmetric Lowe ratio with ratio = conf distance_threshold (matcher.py:101-103)
            ratio = self.conf["model"].get("distance_threshold") if self.mode == "nnr" else None
            p = _mparams(self.mode in ("nnm", "nnr"), None, ratio, self.precision, ratio_mode=1)
            dev = torch.cuda.current_device()
            _lib.check(_lib.lib().sfd2_match_host(_ctx(dev).handle, d0.ctypes.data_as(C.c_void_p), n0,
                                                  d1.ctypes.data_as(C.c_void_p), n1, d, C.byref(p),
                                                  m0.ctypes.data_as(C.c_void_p), s0.ctypes.data_as(C.c_void_p)),
                       "sfd2_match_host")
        return {"matches0": np.maximum(m0, -1).astype(int), "matching_scores0": s0}
