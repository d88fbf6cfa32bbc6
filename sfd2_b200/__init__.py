"""sfd2_b200 - B200 (sm_100a) implementation of the SFD2 extract + match hot path.

Host-side mirror of the reference's Python interface for this path
(feixue94/sfd2: nets/extractor.py, extract_localization.get_model,
hloc/matchers/nearest_neighbor.py, it_loc/matcher.py) over the C ABI of
libsfd2_b200.so (include/sfd2_b200.h).  There is no CPU fallback: every compute
call goes through the CUDA library and raises if it is missing.
"""
from .extractor import ResSegNetV2, get_model, extract_resnet_return, Extractor  # noqa: F401
from .matchers import NearestNeighbor, Matcher, BaseModel, confs as matcher_confs  # noqa: F401

__version__ = "0.1.0"
