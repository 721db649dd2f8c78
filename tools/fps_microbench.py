"""FPS kernels alone at the benchmark shapes: plain (register-resident) vs culled (scene grid).
    python tools/fps_microbench.py            (SAD_B200_LIB=.../libsad_prof.so prints per-phase cycles)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200  # noqa
from sad_b200 import ops
from sad_b200.scenes import make_scenes


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


full = torch.from_numpy(make_scenes(8, 40000, "surface")[0]).cuda()
cur = full
for (N, npnt) in [(40000, 2048), (2048, 1024), (1024, 512), (512, 256)]:
    x = cur.contiguous()
    ops.GRID_MIN_POINTS = 1 << 30
    plain = t(lambda: ops.furthest_point_sample(x, npnt))
    g = ops.build_scene_grid(x)
    build = t(lambda: ops.build_scene_grid(x))
    cull = t(lambda: ops.furthest_point_sample(x, npnt, g), it=3 if os.environ.get("SAD_B200_LIB") else 10)
    one = t(lambda: ops.furthest_point_sample(x, npnt, g, "throughput"), it=3)
    print(f"N={N:6d} -> {npnt:5d}: plain {1e3 * plain:8.1f} us   grid build {1e3 * build:6.1f} us   culled {1e3 * cull:8.1f} us   "
          f"single-CTA culled {1e3 * one:8.1f} us", flush=True)
    inds = ops.furthest_point_sample(x, npnt, g).long()
    cur = torch.gather(x, 1, inds[..., None].expand(-1, -1, 3))
