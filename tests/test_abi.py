"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the
header declares, and the host mirror fails loudly without a GPU (no compute here)."""
import os
import re

import numpy as np
import pytest
import torch

from sfd2_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "sfd2_b200.h")).read()
    declared = set(re.findall(r"SFD2_API\s+[\w\s\*]+?\b(sfd2_\w+)\s*\(", hdr))
    assert declared, "no prototypes parsed from the header"
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    h = _lib.lib()      # raises if the .so is missing or a symbol is absent
    for name in declared:
        assert hasattr(h, name)
    assert h.sfd2_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_lib.ExtractParams) == 32
    assert C.sizeof(_lib.MatchParams) == 24
    assert C.sizeof(_lib.DescSet) == 32


def test_create_rejects_bad_blob_without_touching_cuda():
    import ctypes as C
    h = C.c_void_p()
    buf = C.create_string_buffer(b"NOTAWEIGHTBLOB!!" * 4)
    rc = _lib.lib().sfd2_create(buf, 64, 0, C.byref(h))
    assert rc == -3 and not h
    assert b"magic" in _lib.lib().sfd2_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from sfd2_b200 import get_model, extract_resnet_return, Matcher, matcher_confs
    model, extractor = get_model("ressegnetv2", os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz"), True)
    assert extractor is extract_resnet_return
    with pytest.raises(_lib.Sfd2Error):
        extractor(model, img=torch.zeros(1, 3, 32, 32), topK=10, conf_th=0.001, scales=[1.0])
    with pytest.raises(Exception):
        Matcher(matcher_confs["NNM"])({"descriptors0": np.zeros((4, 128)), "descriptors1": np.zeros((4, 128))})


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "sfd2_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "/root/reference" not in src, f
