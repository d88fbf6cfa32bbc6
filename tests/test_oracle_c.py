"""The plain-C oracle (oracle/nms_oracle.c) against the reference-generated fixtures and the
PyTorch restatement: two independent restatements of the compare-only stages must agree bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import sfd2_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(HERE, "..", "oracle")


@pytest.fixture(scope="module")
def clib():
    subprocess.run(["make", "-s", "-C", ORACLE], check=True)
    lib = C.CDLL(os.path.join(ORACLE, "_build", "liboracle.so"))
    lib.sfd2o_select.restype = C.c_int
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_c_nms_matches_reference_fixtures(clib, golden):
    g = golden("nms_cases")
    for k in [f[3:] for f in g.files if f.startswith("in_")]:
        s = np.ascontiguousarray(g["in_" + k])
        out = np.empty_like(s)
        clib.sfd2o_nms(_p(s), s.shape[0], s.shape[1], 4, _p(out))
        assert np.array_equal(out, g["out_" + k]), k


@pytest.mark.parametrize("name", ["small_96x128", "odd_100x141"])
def test_c_select_matches_reference(clib, golden, name):
    g = golden(name)
    nms = np.ascontiguousarray(g["nms"])
    H, W = nms.shape
    K = int(g["K"])
    xy = np.zeros((K, 2), np.int32)
    sc = np.zeros(K, np.float32)
    n = clib.sfd2o_select(_p(nms), H, W, C.c_float(0.001), 4, K, _p(xy), _p(sc))
    assert n == len(g["scores"])
    assert np.array_equal(xy[:n].astype(np.int16), g["kp_xy"]) and np.array_equal(sc[:n], g["scores"])
    rx, ry, rs = orc.select_keypoints(torch.from_numpy(nms), 0.001, 4, K)
    assert np.array_equal(xy[:n, 0], rx) and np.array_equal(xy[:n, 1], ry)


def test_c_mutual_nn_matches_reference(clib, golden):
    g = golden("match_cases")
    for tag in ["sq", "wide", "tall", "one", "col"]:
        d0, d1 = np.ascontiguousarray(g[f"{tag}_d0"]), np.ascontiguousarray(g[f"{tag}_d1"])
        m = np.zeros(len(d0), np.int32)
        s = np.zeros(len(d0), np.float32)
        clib.sfd2o_mutual_nn(_p(d0), len(d0), _p(d1), len(d1), d0.shape[1], 1, _p(m), _p(s))
        assert np.array_equal(m, g[f"{tag}_hloc_m0"]), tag
        np.testing.assert_allclose((s + 1) / 2, g[f"{tag}_hloc_s0"], atol=1e-6)
