"""CTA-0 timeline of the fused MLP kernel (needs the -DSAD_MLP_PROFILE build: SAD_B200_LIB=.../libsad_prof.so)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200  # noqa
from sad_b200 import mlp as M, _lib

lib = _lib.load()
dump = lib.sad_mlp_profile_dump
dump.argtypes = [ctypes.c_void_p, ctypes.c_void_p]


def layers(ch):
    g = torch.Generator().manual_seed(0)
    return M.prepare_layers([((torch.randn(co, ci, generator=g) / ci ** 0.5).cuda(), torch.zeros(co).cuda())
                             for ci, co in zip(ch[:-1], ch[1:])])


def timeline(tag, fn, first=0, count=60):
    fn(); torch.cuda.synchronize()
    log = np.zeros((4, 2048), dtype=np.int64); n = np.zeros(4, dtype=np.int32)
    dump(log.ctypes.data, n.ctypes.data)          # reset (counters live in registers: nothing to clear on the device)
    fn(); torch.cuda.synchronize()
    dump(log.ctypes.data, n.ctypes.data)
    ev = []
    for role, name in enumerate(["epi", "gat", "mma", "prd"]):
        for i in range(n[role]):
            ev.append((int(log[role, 2 * i + 1]), name, int(log[role, 2 * i])))
    ev.sort()
    t0 = ev[0][0]
    print(f"== {tag}: {len(ev)} events, CTA-0 span {(ev[-1][0] - t0) / 1.9e3:.1f} us (at 1.9 GHz)")
    for (t, name, e) in ev[first:first + count]:
        kind = {1: "wait ", 2: "go   ", 3: "done ", 4: "ready", 5: "stord", 6: "publ "}[e // 100]
        print(f"  {(t - t0):9d} cyc  {name} {kind} layer {e % 100 // 10} ctx {e % 10}")


dev = "cuda"
if len(sys.argv) > 2 and sys.argv[2] == "real":
    # SA1 of the benchmarked step on real (synthetic-scene) data: 8 x 40k surface scenes, FPS + ball query indices
    from sad_b200 import ops
    from sad_b200.scenes import make_scenes
    xyz_np, feat_np = make_scenes(8, 40000, "surface")
    xyz, feat = torch.from_numpy(xyz_np).to(dev), torch.from_numpy(feat_np).to(dev)
    grid = ops.build_scene_grid(xyz)
    inds = ops.furthest_point_sample(xyz, 2048, grid)
    new_xyz = ops.gather_operation(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    idx = ops.ball_query(0.2, 64, xyz, new_xyz, grid)
    m = layers([4, 64, 64, 128])
    timeline("SA1 real data", lambda: M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.2, m), first=int(sys.argv[1]), count=70)
    sys.exit(0)
shapes = [(40000, 2048 * 8, 64, 0, [64, 64, 128]), (2048, 1024 * 8, 32, 128, [128, 128, 256]),
          (512, 256 * 8, 16, 256, [128, 128, 256])]
if len(sys.argv) > 2:
    shapes = [shapes[int(sys.argv[2])]]
for (N, P, S, C, hid) in shapes:
    xyz = torch.rand(1, N, 3, device=dev)
    new_xyz = torch.rand(1, P, 3, device=dev)
    idx = torch.randint(0, N, (1, P, S), device=dev, dtype=torch.int32)
    feat = torch.randn(1, max(C, 1), N, device=dev)
    feat._sad_cl = M.to_cl_bf16(feat) if C else None
    m = layers([max(C, 1) + 3] + hid)
    timeline(f"SA N={N} P={P} S={S} C={C}", lambda: M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.3, m), first=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
x = torch.randn(8, 512, 512, device=dev)
m2 = layers([512, 256, 256])
timeline("pointwise 4096 rows 512->256->256", lambda: M.pointwise_mlp(x, m2), count=80)
