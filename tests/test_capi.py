"""C-ABI checks that need no GPU: the library loads, exports every symbol that
include/sad_ops.h declares, and argument validation returns error codes (never
aborts) before any CUDA work is attempted."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from sad_b200 import _lib
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sad_ops.h")).read()
    return sorted(set(re.findall(r"SAD_API[^;(]*?\b(sad_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for name in ["sad_version", "sad_last_error_string", "sad_furthest_point_sample_fwd",
                 "sad_gather_operation_fwd", "sad_gather_operation_bwd", "sad_ball_query_fwd",
                 "sad_ball_query_adaptive_fwd", "sad_grouping_operation_fwd", "sad_grouping_operation_bwd",
                 "sad_three_nn_fwd", "sad_three_interpolate_fwd", "sad_three_interpolate_bwd"]:
        assert name in syms


def test_library_exports_every_declared_symbol(lib):
    from sad_b200 import _lib
    raw = ctypes.CDLL(_lib.SO_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} declared in sad_ops.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert lib.sad_version() == 1


def test_argument_validation_returns_codes_without_a_gpu(lib):
    # all of these are rejected on the host before any launch
    assert lib.sad_furthest_point_sample_fwd(1, 0, 4, None, None, None) == -1
    assert b"bad sizes" in lib.sad_last_error_string()
    assert lib.sad_furthest_point_sample_fwd(1, 16, 4, None, None, None) == -1
    assert b"null" in lib.sad_last_error_string()
    assert lib.sad_three_nn_fwd(1, 4, 2, None, None, None, None, None) == -1
    assert b"m >= 3" in lib.sad_last_error_string()
    assert lib.sad_ball_query_fwd(1, 16, 4, 0.5, 0, None, None, None, None) == -1
    assert lib.sad_furthest_point_sample_fwd(1, 300000, 4, ctypes.c_void_p(16), ctypes.c_void_p(16), None) == -3
    # the executor's submit call validates its descriptor before touching the CUDA runtime
    from sad_b200 import _lib
    assert lib.sad_engine_submit(None, None) == -1
    assert b"null descriptor" in lib.sad_last_error_string()
    d = _lib.SubmitDesc()
    d.graph_exec, d.n_in = 16, _lib.SUBMIT_MAX_COPIES + 1
    assert lib.sad_engine_submit(ctypes.byref(d), None) == -1
    assert b"bad copy counts" in lib.sad_last_error_string()
    assert ctypes.sizeof(_lib.SubmitDesc) == 3 * 8 + 2 * 4 + 6 * _lib.SUBMIT_MAX_COPIES * 8      # layout of sad_submit_desc
    # empty batches are a no-op success
    assert lib.sad_furthest_point_sample_fwd(0, 16, 4, None, None, None) == 0
    assert lib.sad_grouping_operation_fwd(0, 4, 16, 4, 4, None, None, None, None) == 0


def test_ops_refuse_cpu_tensors():
    import torch
    import sad_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sad_b200.three_nn(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sad_b200.grouping_operation(torch.zeros(1, 4, 8), torch.zeros(1, 2, 2, dtype=torch.int32))
