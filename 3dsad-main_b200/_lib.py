"""ctypes loader for libsad_b200.so (the C ABI in include/sad_ops.h).

Fails loudly: there is no CPU or PyTorch fallback behind these entry points."""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SAD_B200_LIB") or os.path.join(_HERE, "lib", "libsad_b200.so")   # env override: tools only (profiling builds)

_c_int = ctypes.c_int
_c_float = ctypes.c_float
_vp = ctypes.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "sad_version": [],
    "sad_last_error_string": [],
    "sad_furthest_point_sample_cs_fwd": [_c_int, _c_int, _c_int, _vp, _vp, _c_int, _vp],
    "sad_launch_count": [],
    "sad_furthest_point_sample_fwd": [_c_int, _c_int, _c_int, _vp, _vp, _vp],
    "sad_furthest_point_sample_prefix_fwd": [_c_int, _c_int, _c_int, _vp, _vp, _vp, _vp],
    "sad_gather_operation_fwd": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp],
    "sad_gather_operation_bwd": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp],
    "sad_ball_query_fwd": [_c_int, _c_int, _c_int, _c_float, _c_int, _vp, _vp, _vp, _vp],
    "sad_ball_query_adaptive_fwd": [_c_int, _c_int, _c_int, _vp, _c_int, _vp, _vp, _vp, _vp],
    "sad_grouping_operation_fwd": [_c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp],
    "sad_grouping_operation_bwd": [_c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp],
    "sad_three_nn_fwd": [_c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp],
    "sad_three_nn_weights_fwd": [_c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _vp],
    "sad_size_to_radius": [ctypes.c_longlong, _vp, _c_float, _c_float, _c_float, _vp, _vp],
    "sad_gather_points_fwd": [_c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp],
    "sad_three_interpolate_fwd": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp],
    "sad_three_interpolate_bwd": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp],
    "sad_scene_grid_workspace_bytes": [_c_int, _c_int],
    "sad_scene_grid_build": [_c_int, _c_int, _vp, _vp, _vp],
    "sad_ball_query_grid_fwd": [_c_int, _c_int, _c_int, _c_float, _vp, _c_int, _vp, _vp, _vp, _vp, _vp],
    "sad_fps_grid_max_points": [],
    "sad_furthest_point_sample_grid_fwd": [_c_int, _c_int, _c_int, _vp, _vp, _vp, _vp],
    "sad_furthest_point_sample_grid_policy_fwd": [_c_int, _c_int, _c_int, _vp, _vp, _vp, _c_int, _c_int, _vp],
    "sad_mlp_weight_image_bytes": [_c_int, _c_int, _c_int],
    "sad_mlp_pack_weights": [_vp, _c_int, _c_int, _vp, _c_int, _c_int, _vp],
    "sad_shared_mlp_fwd": [_c_int, _c_int, _c_int, _c_int, _vp, _c_int, _vp, _c_int, _vp, _vp, _vp, _c_float, _vp,
                           _c_int, _vp, _c_int, _c_int, _vp, _vp, _vp, _c_int, _vp, _vp, _vp, _vp, _vp],
    "sad_sa_mlp_query": [_c_int] * 8,
    "sad_sa_mlp_instance_info": [_c_int, _vp],
    "sad_sa_mlp_image_bytes": [_c_int],
    "sad_sa_mlp_pack": [_c_int, _vp, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _c_int, _vp],
    "sad_pack_xyzw": [_c_int, _c_int, _vp, _vp, _vp, _vp],
    "sad_sa_mlp_fwd": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _c_float, _vp, _c_int, _vp, _c_int, _vp, _vp,
                       _c_int, _vp, _vp, _vp, _c_int, _vp],
    "sad_sa_mlp_dedup_workspace_bytes": [_c_int, _c_int],
    "sad_sa_mlp_dedup_fwd": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _c_float, _vp, _c_int, _vp, _c_int, _vp, _vp,
                             _c_int, _vp, _vp, _vp, _vp, _c_int, _c_int, _vp],
    "sad_scatter_plan_build": [_c_int, _c_int, ctypes.c_longlong, _vp, _vp, _vp, _vp],
    "sad_interp_plan_build": [_c_int, _c_int, _c_int, _vp, _vp, _vp, _vp],
    "sad_scatter_add_det": [_c_int, _c_int, _c_int, ctypes.c_longlong, _vp, _vp, _vp, _vp, _vp],
    "sad_three_interpolate_bwd_det": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _vp],
    "sad_mlp_tf32_image_bytes": [_c_int, _c_int, _c_int],
    "sad_mlp_tf32_pack": [_vp, _c_int, _c_int, _c_int, _vp],
    "sad_mlp_tf32_fwd": [_c_int, _c_int, _c_int, _c_int, _vp, _c_int, _c_int, _vp, _vp, _vp, _c_int, _vp, _vp, _vp, _c_float, _vp,
                         _c_int, _vp, _c_int, _c_int, _vp, _vp, _vp, _c_int, _vp, _vp, _vp],
    "sad_pw_mlp_image_bytes": [_c_int],
    "sad_pw_mlp_pack": [_c_int, _vp, _vp, _vp, _c_int, _vp],
    "sad_pw_mlp_fwd": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int, _vp, _vp, _vp, _vp, _vp,
                       _c_int, _vp],
    "sad_three_interpolate_cl_fwd": [_c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp],
    "sad_cf_to_cl_bf16": [_c_int, _c_int, _c_int, _vp, _vp, _vp],
    "sad_engine_submit": [_vp, _vp],
}
_RESTYPES = {"sad_last_error_string": ctypes.c_char_p,
             "sad_launch_count": ctypes.c_ulonglong, "sad_scene_grid_workspace_bytes": ctypes.c_longlong, "sad_mlp_weight_image_bytes": ctypes.c_longlong,
             "sad_sa_mlp_image_bytes": ctypes.c_longlong, "sad_pw_mlp_image_bytes": ctypes.c_longlong,
             "sad_mlp_tf32_image_bytes": ctypes.c_longlong, "sad_sa_mlp_dedup_workspace_bytes": ctypes.c_longlong}

SUBMIT_MAX_COPIES = 4


class SubmitDesc(ctypes.Structure):
    """include/sad_ops.h sad_submit_desc"""
    _fields_ = [("graph_exec", _vp), ("wait_event", _vp), ("done_event", _vp),
                ("n_in", _c_int), ("n_out", _c_int),
                ("in_dst", _vp * SUBMIT_MAX_COPIES), ("in_src", _vp * SUBMIT_MAX_COPIES),
                ("in_bytes", ctypes.c_size_t * SUBMIT_MAX_COPIES),
                ("out_dst", _vp * SUBMIT_MAX_COPIES), ("out_src", _vp * SUBMIT_MAX_COPIES),
                ("out_bytes", ctypes.c_size_t * SUBMIT_MAX_COPIES)]


class MlpOpts(ctypes.Structure):
    """include/sad_ops.h sad_mlp_opts"""
    _fields_ = [("tiles_per_cta", ctypes.c_int), ("super_tiles", ctypes.c_int)]


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


class CallProfiler:
    """Context manager: CUDA-event timing of every C-ABI call made inside it (used by
    bench.py / tools for the per-kernel view; never active on the timed path).

        with CallProfiler() as prof: model(...)
        prof.summary() -> {entry_point: {"calls": n, "ms": total_device_ms}}
    """

    def __init__(self, repeat=None):
        """repeat: {entry_point: R} -- issue that (idempotent) call R times back to back inside ONE event pair and
        record elapsed / R: the event pair's own cost (several microseconds around a 148-CTA launch) then does not
        count as kernel time."""
        self.records = []
        self._saved = {}
        self._repeat = dict(repeat or {})

    def __enter__(self):
        import torch
        lib = load()
        for name in SIGNATURES:
            if not (name.endswith("_fwd") or name.endswith("_bwd") or name in ("sad_cf_to_cl_bf16", "sad_scene_grid_build")):
                continue
            fn = getattr(lib, name)
            self._saved[name] = fn

            def wrapped(*args, _fn=fn, _name=name, _reps=int(self._repeat.get(name, 1))):
                a = torch.cuda.Event(enable_timing=True)
                b = torch.cuda.Event(enable_timing=True)
                a.record()
                rc = _fn(*args)
                for _ in range(_reps - 1):
                    if rc == 0:
                        rc = _fn(*args)
                b.record()
                self.records.append((_name, args, a, b, _reps))
                return rc

            setattr(lib, name, wrapped)
        return self

    def __exit__(self, *exc):
        import torch
        lib = load()
        for name, fn in self._saved.items():
            setattr(lib, name, fn)
        torch.cuda.synchronize()
        return False

    def rows(self):
        return [(name, args, a.elapsed_time(b) / reps) for (name, args, a, b, reps) in self.records]

    def summary(self):
        out = {}
        for name, _, ms in self.rows():
            d = out.setdefault(name, {"calls": 0, "ms": 0.0})
            d["calls"] += 1
            d["ms"] += ms
        return out
