"""Build a tools-only variant of the library with extra -D flags (profiling / ablation builds):
    python tools/build_variant.py ablate -DSAD_TOOLS_ABLATE      -> 3dsad-main_b200/lib/libsad_ablate.so
    SAD_B200_LIB=3dsad-main_b200/lib/libsad_ablate.so SAD_ABLATE=1 python bench.py --no-cpu
The product library never carries these flags."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "3dsad-main_b200", "csrc")
sys.path.insert(0, CSRC)
import build as B  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
objdir = os.path.join(B.PKG, "build", name)
os.makedirs(objdir, exist_ok=True)
objs = []


def one(src):
    obj = os.path.join(objdir, src.replace(".cu", ".o"))
    subprocess.run([B.NVCC, *B.FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj], check=True)
    return obj


with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(one, B.SOURCES))
so = os.path.join(B.LIB_DIR, f"libsad_{name}.so")
subprocess.run([B.NVCC, "-shared", "-o", so, *objs, "-cudart", "static"], check=True)
print(so)
