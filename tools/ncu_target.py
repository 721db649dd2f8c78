"""Small driver for `ncu`: each fused-MLP kernel of the step a few times at the benchmark's shapes (real scene data for
SA1 / SA2).  tools/gpu_ncu.sh runs it plain first, then under ncu --set full."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200  # noqa
from sad_b200 import mlp as M, ops
from sad_b200.scenes import make_scenes

dev, B = "cuda", 8
g = torch.Generator().manual_seed(0)


def layers(ch):
    return M.prepare_layers([((torch.randn(co, ci, generator=g) / ci ** 0.5).cuda(), 0.1 * torch.randn(co, generator=g).cuda())
                             for ci, co in zip(ch[:-1], ch[1:])])


xyz_np, feat_np = make_scenes(B, 40000, "surface")
xyz, feat = torch.from_numpy(xyz_np).to(dev), torch.from_numpy(feat_np).to(dev)
grid = ops.build_scene_grid(xyz)
inds = ops.furthest_point_sample(xyz, 2048, grid, "throughput")
x1 = ops.gather_points(xyz, inds)
idx1 = ops.ball_query(0.2, 64, xyz, x1, grid)
m1, m2, m3 = layers([4, 64, 64, 128]), layers([131, 128, 128, 256]), layers([259, 128, 128, 256])
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    f1 = M.sa_group_mlp(xyz, x1, feat, idx1, 0.2, m1)
i2 = ops.furthest_point_sample(x1, 1024)
x2 = ops.gather_points(x1, i2)
idx2 = ops.ball_query(0.4, 32, x1, x2)
for _ in range(reps):
    f2 = M.sa_group_mlp(x1, x2, f1, idx2, 0.4, m2)
i3 = ops.furthest_point_sample(x2, 512)
x3 = ops.gather_points(x2, i3)
idx3 = ops.ball_query(0.8, 16, x2, x3)
for _ in range(reps):
    f3 = M.sa_group_mlp(x2, x3, f2, idx3, 0.8, m3)
_, nn_i, nn_w = ops.three_nn_weights(x2, x3)
mf = layers([512, 256, 256])
for _ in range(reps):
    fp = M.fp_interp_mlp(f3, f2, nn_i, nn_w, mf)
mv = layers([256, 256, 256, 259])
for _ in range(reps):
    vx, vf = M.vote_mlp_fast(x2, fp, mv)
# tf32 mode of the SA2 stage, and the drop-in three_interpolate at the FP2 shape with B = 256 (working set > L2)
m2t = M.prepare_layers(m2.layers, dtype="tf32")
for _ in range(reps):
    M.sa_group_mlp(x1, x2, f1, idx2, 0.4, m2t)
Bi, n, m = 256, 1024, 512
u, k = torch.rand(Bi, n, 3, device=dev) * 6, torch.rand(Bi, m, 3, device=dev) * 6
_, ii, ww = ops.three_nn_weights(u, k)
ff = torch.randn(Bi, 256, m, device=dev)
for _ in range(reps):
    ops.three_interpolate(ff, ii, ww)
torch.cuda.synchronize()
print("ok")
