"""Where the duplicate-free SA launch spends its time: the plain instance, the 16-sample sibling on a pre-trimmed idx
(no plan, no indirection), and the planned launch, on real SA1 / SA2 inputs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sad_b200  # noqa
from sad_b200 import mlp as M, ops
from sad_b200.scenes import make_scenes
from stage_bench import t, layers

dev, B = "cuda", 8
xyz_np, feat_np = make_scenes(B, 40000, "surface")
xyz, feat = torch.from_numpy(xyz_np).to(dev), torch.from_numpy(feat_np).to(dev)
grid = ops.build_scene_grid(xyz)
inds = ops.furthest_point_sample(xyz, 2048, grid)
new_xyz = ops.gather_points(xyz, inds)
idx = ops.ball_query(0.2, 64, xyz, new_xyz, grid)
m1 = layers([4, 64, 64, 128])
M.prepack_xyzw(xyz, feat)
for name, fn in (("sa1 plain S=64", lambda: M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.2, m1)),):
    M.DEDUP_SA[0] = False
    print(name, round(t(fn)[0], 1), flush=True)
M.DEDUP_SA[0] = True
print("sa1 dedup", round(t(lambda: M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.2, m1))[0], 1), flush=True)
M.DEDUP_SA[0] = False
for S_ in (32, 16):
    idt = idx[:, :, :S_].contiguous()
    print(f"sa1 trimmed idx S={S_} (no plan; inexact)", round(t(lambda: M.sa_group_mlp(xyz, new_xyz, feat, idt, 0.2, m1))[0], 1), flush=True)
f1 = M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.2, m1)
inds2 = ops.furthest_point_sample(new_xyz, 1024)
x2 = ops.gather_points(new_xyz, inds2)
idx2 = ops.ball_query(0.4, 32, new_xyz, x2)
m2 = layers([131, 128, 128, 256])
M.FAST_SA[0] = "single"
print("sa2 plain S=32", round(t(lambda: M.sa_group_mlp(new_xyz, x2, f1, idx2, 0.4, m2))[0], 1), flush=True)
M.DEDUP_SA[0] = True
print("sa2 dedup", round(t(lambda: M.sa_group_mlp(new_xyz, x2, f1, idx2, 0.4, m2))[0], 1), flush=True)
M.DEDUP_SA[0] = False
id16 = idx2[:, :, :16].contiguous()
print("sa2 trimmed idx S=16 (no plan; inexact)", round(t(lambda: M.sa_group_mlp(new_xyz, x2, f1, id16, 0.4, m2))[0], 1), flush=True)
# the plan kernel alone
import ctypes
lib = sad_b200._lib.load()
vp = ctypes.c_void_p
