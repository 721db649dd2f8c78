"""ctypes loader for libsad_b200.so (the C ABI in include/sad_ops.h).

Fails loudly: there is no CPU or PyTorch fallback behind these entry points."""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SAD_B200_LIB", os.path.join(_HERE, "lib", "libsad_b200.so"))

_c_int = ctypes.c_int
_c_float = ctypes.c_float
_vp = ctypes.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "sad_version": [],
    "sad_last_error_string": [],
    "sad_fps_force_cluster_size": [_c_int],
    "sad_furthest_point_sample_fwd": [_c_int, _c_int, _c_int, _vp, _vp, _vp],
    "sad_gather_operation_fwd": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp],
    "sad_gather_operation_bwd": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp],
    "sad_ball_query_fwd": [_c_int, _c_int, _c_int, _c_float, _c_int, _vp, _vp, _vp, _vp],
    "sad_ball_query_adaptive_fwd": [_c_int, _c_int, _c_int, _vp, _c_int, _vp, _vp, _vp, _vp],
    "sad_grouping_operation_fwd": [_c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp],
    "sad_grouping_operation_bwd": [_c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp],
    "sad_three_nn_fwd": [_c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp],
    "sad_three_interpolate_fwd": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp],
    "sad_three_interpolate_bwd": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp],
}
_RESTYPES = {"sad_last_error_string": ctypes.c_char_p, "sad_fps_force_cluster_size": None}

_lib = None
_lock = threading.Lock()


class SadLibraryError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises if the extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(SO_PATH):
                raise SadLibraryError(
                    f"libsad_b200.so not found at {SO_PATH}: build it with "
                    "`python 3dsad-main_b200/csrc/build.py` (or __graft_entry__.build()). "
                    "There is no CPU fallback.")
            lib = ctypes.CDLL(SO_PATH)
            for name, argtypes in SIGNATURES.items():
                fn = getattr(lib, name)      # AttributeError if the .so lacks a declared symbol
                fn.argtypes = argtypes
                fn.restype = _RESTYPES.get(name, _c_int)
            if lib.sad_version() != 1:
                raise SadLibraryError(f"libsad_b200 ABI {lib.sad_version()} != 1")
            _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().sad_last_error_string()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
