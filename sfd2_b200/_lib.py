"""ctypes binding of libsfd2_b200.so (include/sfd2_b200.h).  Fails loudly: no fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsfd2_b200.so")

PREC = {"fp32": 0, "exact": 1, "fast": 2, "mixed": 3}
IMG_F32_NCHW, IMG_U8_NHWC = 0, 1
DESC_ROWS, DESC_COLS = 0, 1
DESC_DIM = 128
ABI_VERSION = 2


class ExtractParams(C.Structure):
    _fields_ = [("conf_th", C.c_float), ("nms_radius", C.c_int32), ("border", C.c_int32),
                ("topk", C.c_int32), ("precision", C.c_int32), ("use_stability", C.c_int32),
                ("border_w", C.c_int32), ("border_h", C.c_int32)]


class MatchParams(C.Structure):
    _fields_ = [("do_mutual_check", C.c_int32), ("distance_threshold", C.c_float),
                ("ratio_threshold", C.c_float), ("precision", C.c_int32), ("ratio_mode", C.c_int32),
                ("layout", C.c_int32)]


class DescSet(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_int32), ("layout", C.c_int32), ("count", C.c_void_p),
                ("ids", C.c_void_p)]


class Sfd2Error(RuntimeError):
    pass


def device_index(device=None) -> int:
    """CUDA device index of `device` (None / 'cuda' / torch.device without an index = the CURRENT device,
    not GPU 0: under torchrun every rank has set its own)."""
    import torch
    if device is None:
        return torch.cuda.current_device()
    if isinstance(device, int):
        return device
    d = torch.device(device)
    if d.type != "cuda":
        raise Sfd2Error(f"sfd2_b200 runs on CUDA devices only (got {d})")
    return torch.cuda.current_device() if d.index is None else d.index


_lib = None

_PROTOS = {
    "sfd2_abi_version": (C.c_int, []),
    "sfd2_last_error": (C.c_char_p, []),
    "sfd2_create": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "sfd2_destroy": (C.c_int, [C.c_void_p]),
    "sfd2_extract_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(ExtractParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "sfd2_extract_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(ExtractParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfd2_extract_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sfd2_preprocess_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "sfd2_match_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                 C.POINTER(MatchParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfd2_match_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                  C.POINTER(MatchParams), C.c_void_p, C.c_void_p]),
    "sfd2_match_pairs_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.POINTER(MatchParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfd2_match_batched_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.POINTER(MatchParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfd2_match_one_to_many_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                             C.POINTER(MatchParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfd2_debug_fetch": (C.c_longlong, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]),
    "sfd2_launch_count": (C.c_longlong, [C.c_void_p]),
    "sfd2_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "sfd2_profile_read": (C.c_longlong, [C.c_void_p, C.c_char_p, C.c_longlong]),
    "sfd2_nms_select_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(ExtractParams),
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfd2_debug_conv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}
EXPORTS = tuple(_PROTOS)


def lib():
    """The loaded library.  Raises Sfd2Error if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Sfd2Error(f"{LIB_PATH} is missing: build it with `python -m sfd2_b200.build` "
                            "(there is no CPU fallback)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(h, name)     # AttributeError if the .so does not export what the header declares
            fn.restype, fn.argtypes = res, args
        if h.sfd2_abi_version() != ABI_VERSION:
            raise Sfd2Error("libsfd2_b200.so ABI version mismatch")
        _lib = h
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().sfd2_last_error().decode("utf-8", "replace")
        raise Sfd2Error(f"{what} failed ({rc}): {msg}")


class Context:
    """Owns one native sfd2_ctx (weights + workspace) on one CUDA device."""

    def __init__(self, blob, device: int):
        """blob: folded weight blob (bytes), or None for a matcher-only context."""
        self._h = C.c_void_p()
        if blob is None:
            self._blob = None
            check(lib().sfd2_create(None, 0, int(device), C.byref(self._h)), "sfd2_create")
        else:
            self._blob = C.create_string_buffer(blob, len(blob))
            check(lib().sfd2_create(self._blob, len(blob), int(device), C.byref(self._h)), "sfd2_create")
        self.device = int(device)

    @property
    def handle(self):
        if not self._h:
            raise Sfd2Error("context already destroyed")
        return self._h

    def launch_count(self) -> int:
        return int(lib().sfd2_launch_count(self.handle))

    def profile(self, enable: bool):
        check(lib().sfd2_profile(self.handle, int(enable)), "sfd2_profile")

    def profile_read(self) -> dict:
        """{label: (launches, total_ms)} since the last read (synchronises the device)."""
        buf = C.create_string_buffer(1 << 16)
        n = lib().sfd2_profile_read(self.handle, buf, len(buf))
        if n < 0:
            check(int(n), "sfd2_profile_read")
        out = {}
        for line in buf.value.decode().splitlines():
            k, cnt, ms = line.split("\t")
            out[k] = (int(cnt), float(ms))
        return out

    def close(self):
        if self._h:
            lib().sfd2_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
