"""Culled FPS at the headline shape (8 x 40k -> 2048) for every legal cluster size: latency and SM*time.
    python tools/fps_cs_compare.py [N] [npoint]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200  # noqa
from sad_b200 import ops, _lib
from sad_b200.scenes import make_scenes

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
npnt = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
x = torch.from_numpy(make_scenes(8, N, "surface")[0]).cuda()
g = ops.build_scene_grid(x)
lib = _lib.load()


def t(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


ref = None
for cs in (0, -1, 4, 8):
    try:
        ms = t(lambda: ops.furthest_point_sample(x, npnt, g, "latency", False, cs))
    except RuntimeError as e:
        print(f"cs={cs}: {e}")
        continue
    out = ops.furthest_point_sample(x, npnt, g, "latency", False, cs)
    if ref is None:
        ref = out
    same = bool((out == ref).all())
    eff = cs if cs else "auto"
    print(f"N={N} npoint={npnt} cs={eff}: {1e3 * ms:8.1f} us  ({1e3 * ms / (npnt - 1) * 1.0:6.3f} us/pick)  same={same}", flush=True)
