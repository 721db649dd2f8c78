"""Build libsad_b200.so in-tree with nvcc for sm_100a (no torch headers: pure C ABI).

    python 3dsad-main_b200/csrc/build.py [--force] [--verbose]

The .so lands in 3dsad-main_b200/lib/ (git-ignored, but it travels to the GPU box).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(PKG, "build")
SO = os.path.join(LIB_DIR, "libsad_b200.so")
SOURCES = ["capi.cu", "fps.cu", "fps_cull.cu", "grid.cu", "search.cu", "gather.cu", "interp.cu", "mlp.cu", "mlp_sa.cu", "mlp_pw.cu", "mlp_tf32.cu", "scatter.cu", "engine.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
         "-I", os.path.join(ROOT, "include")]


def _deps(src):
    deps = [os.path.join(HERE, src), os.path.join(HERE, "sad_common.cuh"),
            os.path.join(ROOT, "include", "sad_ops.h")]
    deps += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    return deps


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    sources = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    objs, jobs = [], []
    for src in sources:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, _deps(src)):
            cmd = [NVCC, *FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(SO, objs):
        run([NVCC, "-shared", "-o", SO, *objs, "-cudart", "static"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
