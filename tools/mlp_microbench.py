"""Fused MLP kernel alone: fixed cost (1 tile) and per-tile cost at the benchmark's stage shapes."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200  # noqa
from sad_b200 import mlp as M


def t(fn, it=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return 1e3 * ts[len(ts) // 2]


def layers(ch):
    g = torch.Generator().manual_seed(0)
    return M.prepare_layers([((torch.randn(co, ci, generator=g) / ci ** 0.5).cuda(), torch.zeros(co).cuda())
                             for ci, co in zip(ch[:-1], ch[1:])])


dev = "cuda"
for dyn in (True, False):
    M.DYNAMIC_TILES = dyn
    print("dynamic tiles" if dyn else "static tiles")
    for rows in (128, 128 * 148, 128 * 148 * 4):
        x = torch.randn(1, 256, rows, device=dev)
        x._sad_cl = M.to_cl_bf16(x)
        m2, m3 = layers([256, 256, 256]), layers([256, 256, 256, 259])
        print(f"  pointwise rows={rows:6d}: 2-layer {t(lambda: M.pointwise_mlp(x, m2)):7.1f} us   3-layer(259) "
              f"{t(lambda: M.pointwise_mlp(x, m3, last_relu=False, want_cl=False)):7.1f} us")
    for (N, P, S, C, hid) in [(2048, 2, 64, 0, [64, 64, 128]), (40000, 2048 * 8, 64, 0, [64, 64, 128]),
                              (2048, 4, 32, 128, [128, 128, 256]), (2048, 1024 * 8, 32, 128, [128, 128, 256])]:
        B = 1
        xyz = torch.rand(B, N, 3, device=dev)
        new_xyz = torch.rand(B, P, 3, device=dev)
        idx = torch.randint(0, N, (B, P, S), device=dev, dtype=torch.int32)
        feat = torch.randn(B, max(C, 1), N, device=dev)
        m = layers([max(C, 1) + 3] + hid)
        print(f"  SA stage N={N} P={P} S={S} C={C}: {t(lambda: M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.3, m)):7.1f} us "
              f"({P * S // 128} tiles)")
