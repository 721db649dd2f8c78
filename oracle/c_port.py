"""ctypes front-end + build recipe for oracle/sad_oracle_c.c -- TEST INFRASTRUCTURE ONLY.

The C port is the multi-threaded CPU checker / timed CPU baseline ("port" kind in
bench.py's cpu_baseline).  PARITY UNPINNED w.r.t. the reference (README-only mount);
it is verified bit-for-bit against the NumPy oracle in tests/test_oracle.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "sad_oracle_c.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "libsad_oracle.so")

_lib = None


def build(force: bool = False) -> str:
    """gcc -O3 -fopenmp -ffp-contract=off (no FMA contraction: arithmetic contract H1)."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O3", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
               "-o", _SO, _SRC, "-lm"]
        subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def set_threads(n: int) -> None:
    """Use `n` OpenMP threads from now on (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().orc_set_threads(int(n))


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def furthest_point_sample(xyz, npoint):
    xyz = _f(xyz)
    B, N, _ = xyz.shape
    out = np.zeros((B, npoint), dtype=np.int32)
    rc = lib().orc_furthest_point_sample(B, N, npoint, _p(xyz, ctypes.c_float), _p(out, ctypes.c_int32))
    if rc:
        raise ValueError("orc_furthest_point_sample: bad arguments")
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    B, N, _ = xyz.shape
    P = new_xyz.shape[1]
    out = np.zeros((B, P, nsample), dtype=np.int32)
    rc = lib().orc_ball_query(B, N, P, ctypes.c_float(radius), None, nsample, _p(xyz, ctypes.c_float),
                              _p(new_xyz, ctypes.c_float), _p(out, ctypes.c_int32))
    if rc:
        raise ValueError("orc_ball_query: bad arguments")
    return out


def ball_query_adaptive(radius_t, nsample, xyz, new_xyz):
    xyz, new_xyz, radius_t = _f(xyz), _f(new_xyz), _f(radius_t)
    B, N, _ = xyz.shape
    P = new_xyz.shape[1]
    out = np.zeros((B, P, nsample), dtype=np.int32)
    rc = lib().orc_ball_query(B, N, P, ctypes.c_float(0.0), _p(radius_t, ctypes.c_float), nsample,
                              _p(xyz, ctypes.c_float), _p(new_xyz, ctypes.c_float), _p(out, ctypes.c_int32))
    if rc:
        raise ValueError("orc_ball_query: bad arguments")
    return out


def three_nn(unknown, known):
    unknown, known = _f(unknown), _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist = np.zeros((B, n, 3), dtype=np.float32)
    idx = np.zeros((B, n, 3), dtype=np.int32)
    rc = lib().orc_three_nn(B, n, m, _p(unknown, ctypes.c_float), _p(known, ctypes.c_float),
                            _p(dist, ctypes.c_float), _p(idx, ctypes.c_int32))
    if rc:
        raise ValueError("three_nn requires m >= 3 known points")
    return dist, idx


def grouping_operation(features, idx):
    features, idx = _f(features), _i(idx)
    B, C, N = features.shape
    P, S = idx.shape[1], idx.shape[2]
    out = np.zeros((B, C, P, S), dtype=np.float32)
    lib().orc_grouping_operation(B, C, N, P, S, _p(features, ctypes.c_float), _p(idx, ctypes.c_int32),
                                 _p(out, ctypes.c_float))
    return out


def gather_operation(features, idx):
    idx = _i(idx)
    return grouping_operation(features, idx[:, :, None])[..., 0]


def three_interpolate(features, idx, weight):
    features, idx, weight = _f(features), _i(idx), _f(weight)
    B, C, m = features.shape
    n = idx.shape[1]
    out = np.zeros((B, C, n), dtype=np.float32)
    lib().orc_three_interpolate(B, C, m, n, _p(features, ctypes.c_float), _p(idx, ctypes.c_int32),
                                _p(weight, ctypes.c_float), _p(out, ctypes.c_float))
    return out


if __name__ == "__main__":
    print(build(force=True))
