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

__all__ = ["BaseModel", "NearestNeighbor", "NearestNeighborMixin", "Matcher", "confs", "match_dev", "match_batched",
           "match_one_to_many", "match_sets_dev", "match_pairs_dev", "feature_matching"]

_CTX = {}


def _ctx(device_index: int) -> "_lib.Context":
    """The matcher needs no network weights: one matcher-only native context per device (sfd2_create(NULL))."""
    if device_index not in _CTX:
        _CTX[device_index] = _lib.Context(None, device_index)
    return _CTX[device_index]


def _mparams(mutual, dist_th, ratio_th, precision, ratio_mode=0, layout=0):
    return _lib.MatchParams(do_mutual_check=int(bool(mutual)),
                            distance_threshold=float(dist_th) if dist_th else 0.0,
                            ratio_threshold=float(ratio_th) if ratio_th else 0.0,
                            precision=_lib.PREC[precision], ratio_mode=int(ratio_mode), layout=int(layout))


PLAIN_CODES, HLOC_SCORES, I64 = 0x100, 0x200, 0x400     # sfd2_match_params.ratio_mode flag bits (sfd2_b200.h)


def match_dev(d0: torch.Tensor, d1: torch.Tensor, mutual=True, dist_th=None, ratio_th=None, precision="exact",
              ratio_mode=0, raw=False, layout="rows"):
    """d0 [N,D], d1 [M,D] CUDA float32 row-major (layout="cols": [D,N], [D,M] as hloc passes them)
    -> (matches0 int32 [N], sim0 float32 [N]) on the device.
    matches0 < 0 means no match; with raw=True the native codes are kept (-1 = the row failed its own
    ratio/distance test, -2 = it failed only the mutual check)."""
    assert d0.is_cuda and d1.is_cuda and d0.dtype == torch.float32 and d1.dtype == torch.float32
    d0, d1 = d0.contiguous(), d1.contiguous()
    cols = layout == "cols"
    n0, n1 = (d0.shape[1], d1.shape[1]) if cols else (d0.shape[0], d1.shape[0])
    d = d0.shape[0] if cols else d0.shape[1]
    dev = d0.device
    # the finish stage writes every one of the n0 rows (also when d1 is empty), so no fill kernels are needed
    m0 = torch.empty((n0,), dtype=torch.int32, device=dev)
    s0 = torch.empty((n0,), dtype=torch.float32, device=dev)
    if n0 == 0:
        return m0, s0
    # raw=False: the finish stage itself reports every unmatched row as -1 (SFD2_MATCH_PLAIN_CODES), no clamp launch
    p = _mparams(mutual, dist_th, ratio_th, precision, int(ratio_mode) | (0 if raw else PLAIN_CODES), layout=int(cols))
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().sfd2_match_dev(_ctx(_lib.device_index(dev)).handle, d0.data_ptr(), n0, d1.data_ptr(), n1, d,
                                         C.byref(p), m0.data_ptr(), s0.data_ptr(), st), "sfd2_match_dev")
    return m0, s0


def match_sets_dev(sets, pair_a, pair_b, mutual=True, dist_th=None, ratio_th=None, precision="exact", ratio_mode=0,
                   raw=False):
    """The grouped native call (sfd2_match_pairs_dev): ONE launch for all pairs.
    sets: list of dicts {"data": CUDA float32 tensor [n,128] (or [128,n] with "layout": "cols"),
                         "count": optional CUDA int32 tensor (1 element; valid rows live on the device),
                         "ids": optional CUDA int32 tensor [n] (rows with -1 take no part; matches report original rows)}
    pair k matches sets[pair_a[k]] (rows) against sets[pair_b[k]].
    -> (matches0 int32, sim0 float32) flat device tensors; pair k's rows start at sum of n(pair_a[q]) for q < k."""
    dev = sets[0]["data"].device
    arr = (_lib.DescSet * len(sets))()
    keep = []
    ns = []
    for i, s_ in enumerate(sets):
        t = s_["data"]
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == dev
        cols = s_.get("layout", "rows") == "cols"
        n = t.shape[1] if cols else t.shape[0]
        arr[i].data, arr[i].n, arr[i].layout = t.data_ptr(), int(n), int(cols)
        cnt, ids = s_.get("count"), s_.get("ids")
        if cnt is not None:
            assert cnt.is_cuda and cnt.dtype == torch.int32
            arr[i].count = cnt.data_ptr()
        if ids is not None:
            ids = ids.to(torch.int32).contiguous()
            assert ids.is_cuda and ids.numel() == n
            arr[i].ids = ids.data_ptr()
            keep.append(ids)
        ns.append(int(n))
    pa = np.ascontiguousarray(pair_a, np.int32)
    pb = np.ascontiguousarray(pair_b, np.int32)
    total = int(sum(ns[a] for a in pa))
    m0 = torch.empty((total,), dtype=torch.int32, device=dev)
    s0 = torch.empty((total,), dtype=torch.float32, device=dev)
    if len(pa) == 0 or total == 0:
        return m0, s0
    p = _mparams(mutual, dist_th, ratio_th, precision, int(ratio_mode) | (0 if raw else PLAIN_CODES))
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().sfd2_match_pairs_dev(_ctx(_lib.device_index(dev)).handle, arr, len(sets),
                                               pa.ctypes.data_as(C.c_void_p), pb.ctypes.data_as(C.c_void_p), len(pa),
                                               C.byref(p), m0.data_ptr(), s0.data_ptr(), st), "sfd2_match_pairs_dev")
    return m0, s0


def match_pairs_dev(descriptors: torch.Tensor, counts: torch.Tensor, idx0, idx1, mutual=True, dist_th=None,
                    ratio_th=None, precision="exact"):
    """Pairs of images out of ONE extractor result, in one grouped launch, with no host synchronisation:
    descriptors [n,K,128] / counts int32 [n] are sfd2_extract_dev's fixed-capacity outputs (device); pair p matches
    image idx0[p] against image idx1[p], the valid row counts are read by the kernels from `counts`.
    -> matches0 int32 [P,K] (-1 = unmatched, rows beyond counts[idx0[p]] are -1), sim0 float32 [P,K]."""
    n, K, _ = descriptors.shape
    descriptors = descriptors.contiguous()
    counts = counts.contiguous()
    sets = [{"data": descriptors[i], "count": counts[i:i + 1]} for i in range(n)]
    pa = [int(i) for i in (idx0.tolist() if torch.is_tensor(idx0) else idx0)]
    pb = [int(i) for i in (idx1.tolist() if torch.is_tensor(idx1) else idx1)]
    m0, s0 = match_sets_dev(sets, pa, pb, mutual, dist_th, ratio_th, precision)
    return m0.view(len(pa), K), s0.view(len(pa), K)


def match_batched(d0: torch.Tensor, off0, d1: torch.Tensor, off1, mutual=True, dist_th=None, precision="exact",
                  ratio_th=None, ratio_mode=0):
    """Many pairs in one native call; off0/off1 are python/numpy int sequences of length npairs+1."""
    d0, d1 = d0.contiguous(), d1.contiguous()
    o0 = np.ascontiguousarray(off0, np.int32)
    o1 = np.ascontiguousarray(off1, np.int32)
    npairs = len(o0) - 1
    dev = d0.device
    m0 = torch.full((d0.shape[0],), -1, dtype=torch.int32, device=dev)
    s0 = torch.zeros((d0.shape[0],), dtype=torch.float32, device=dev)
    p = _mparams(mutual, dist_th, ratio_th, precision, int(ratio_mode) | PLAIN_CODES)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().sfd2_match_batched_dev(_ctx(_lib.device_index(dev)).handle, d0.data_ptr(),
                                                 o0.ctypes.data_as(C.c_void_p), d1.data_ptr(),
                                                 o1.ctypes.data_as(C.c_void_p), npairs, d0.shape[1], C.byref(p),
                                                 m0.data_ptr(), s0.data_ptr(), st), "sfd2_match_batched_dev")
    return m0, s0


def match_one_to_many(q: torch.Tensor, db: torch.Tensor, db_off, mutual=True, dist_th=None, precision="exact",
                      ratio_th=None, ratio_mode=0, db_ids=None):
    """One query set q [N,128] against ndb db sets stored back to back in db (row offsets db_off, length
    ndb+1) in ONE grouped launch - the localizer's pattern (a query against its retrieved db images,
    it_loc/localize_cv2.py:705-731).  db_ids (optional int tensor, one entry per db row): rows with -1 (keypoints
    without a 3-D point, :540-555) take no part and matches are LOCAL ORIGINAL row indices of the db set (the
    reference's valid_ids remap, :557-559, happens on the device).
    -> (matches0 int32 [ndb, N] or -1, sim0 float32 [ndb, N])."""
    q, db = q.contiguous(), db.contiguous()
    off = np.ascontiguousarray(db_off, np.int64)
    ndb = len(off) - 1
    if precision == "fp32" or db_ids is None:
        dev = q.device
        off32 = off.astype(np.int32)
        m0 = torch.full((ndb, q.shape[0]), -1, dtype=torch.int32, device=dev)
        s0 = torch.zeros((ndb, q.shape[0]), dtype=torch.float32, device=dev)
        if db_ids is not None:
            raise _lib.Sfd2Error("db_ids needs a tcgen05 precision mode")
        p = _mparams(mutual, dist_th, ratio_th, precision, int(ratio_mode) | PLAIN_CODES)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().sfd2_match_one_to_many_dev(_ctx(_lib.device_index(dev)).handle, q.data_ptr(), q.shape[0],
                                                         db.data_ptr(), off32.ctypes.data_as(C.c_void_p), ndb, q.shape[1],
                                                         C.byref(p), m0.data_ptr(), s0.data_ptr(), st),
                   "sfd2_match_one_to_many_dev")
        return m0, s0
    ids = db_ids.to(device=q.device, dtype=torch.int32).contiguous()
    sets = [{"data": q}] + [{"data": db[off[i]:off[i + 1]], "ids": ids[off[i]:off[i + 1]]} for i in range(ndb)]
    m0, s0 = match_sets_dev(sets, [0] * ndb, list(range(1, ndb + 1)), mutual, dist_th, ratio_th, precision, ratio_mode)
    return m0.view(ndb, q.shape[0]), s0.view(ndb, q.shape[0])


def feature_matching(desc_q, desc_db, matcher=None, label_q=None, label_db=None, db_3D_ids=None, precision="exact"):
    """it_loc/localize_cv2.py:511-560 (labels unsupported: outside the hot path).  numpy in, numpy int matches out.
    desc_db / db_3D_ids may be LISTS (one entry per retrieved db image): the whole query-vs-db-images loop of
    match_cluster_2D (:563-649) then runs as one grouped launch and a list of match arrays is returned.
    The db_3D_ids != -1 subset and the index remap (:540-559) happen on the device."""
    if label_q is not None or label_db is not None:
        raise NotImplementedError("label-aware matching (nnml) is outside the hot path")
    many = isinstance(desc_db, (list, tuple))
    dbs = list(desc_db) if many else [desc_db]
    ids = list(db_3D_ids) if many and db_3D_ids is not None else ([db_3D_ids] if db_3D_ids is not None else None)
    dev = torch.device("cuda", torch.cuda.current_device())
    q = torch.from_numpy(np.ascontiguousarray(desc_q, np.float32)).to(dev)
    off = np.concatenate([[0], np.cumsum([len(d) for d in dbs])]).astype(np.int64)
    db = torch.from_numpy(np.ascontiguousarray(np.concatenate(dbs), np.float32)).to(dev)
    idt = None if ids is None else torch.from_numpy(np.concatenate([np.asarray(i) for i in ids]).astype(np.int32)).to(dev)
    mode = matcher.mode if matcher is not None else "nnm"
    ratio = matcher.conf["model"].get("distance_threshold") if (matcher is not None and mode == "nnr") else None
    m0, _ = match_one_to_many(q, db, off, mutual=mode in ("nnm", "nnr"), precision=precision, ratio_th=ratio,
                              ratio_mode=1, db_ids=idt)
    out = [m.astype(int) for m in m0.cpu().numpy()]
    return out if many else out[0]


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
        prec = self.conf.get("precision", "exact")
        B, D, N = d0.shape
        M = d1.shape[2]
        dev = d0.device
        if prec == "fp32":      # CUDA-core reference mode: row-major operands, scores / dtype glue in torch
            ms, ss = [], []
            for b in range(B):
                m0, s0 = match_dev(d0[b].float().t().contiguous(), d1[b].float().t().contiguous(),
                                   self.conf["do_mutual_check"], self.conf["distance_threshold"],
                                   self.conf["ratio_threshold"], prec, ratio_mode=0, raw=True)
                ss.append(torch.where(m0 == -1, s0.new_tensor(0), (s0 + 1) / 2))
                ms.append(m0.clamp(min=-1).long())
            return {"matches0": torch.stack(ms), "matching_scores0": torch.stack(ss)}
        # tcgen05 modes: the kernels read hloc's [D, N] layout directly and write int64 matches and
        # (sim + 1) / 2 scores (find_nn, nearest_neighbor.py:14-15) - no transpose / where / clamp / stack launches
        d0 = d0 if (d0.dtype == torch.float32 and d0.is_contiguous()) else d0.float().contiguous()
        d1 = d1 if (d1.dtype == torch.float32 and d1.is_contiguous()) else d1.float().contiguous()
        m0 = torch.empty((B, N), dtype=torch.int64, device=dev)
        s0 = torch.empty((B, N), dtype=torch.float32, device=dev)
        if N == 0:
            return {"matches0": m0, "matching_scores0": s0}
        arr = (_lib.DescSet * (2 * B))()
        for b in range(B):
            arr[2 * b].data, arr[2 * b].n, arr[2 * b].layout = d0[b].data_ptr(), N, _lib.DESC_COLS
            arr[2 * b + 1].data, arr[2 * b + 1].n, arr[2 * b + 1].layout = d1[b].data_ptr(), M, _lib.DESC_COLS
        pa = np.arange(0, 2 * B, 2, dtype=np.int32)
        pb = pa + 1
        p = _mparams(self.conf["do_mutual_check"], self.conf["distance_threshold"], self.conf["ratio_threshold"], prec,
                     ratio_mode=0 | PLAIN_CODES | HLOC_SCORES | I64)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().sfd2_match_pairs_dev(_ctx(_lib.device_index(dev)).handle, arr, 2 * B,
                                                   pa.ctypes.data_as(C.c_void_p), pb.ctypes.data_as(C.c_void_p), B,
                                                   C.byref(p), m0.data_ptr(), s0.data_ptr(), st), "sfd2_match_pairs_dev")
        return {"matches0": m0, "matching_scores0": s0}


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
