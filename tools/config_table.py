"""BASELINE.json configs[0] and configs[2], GPU next to the CPU path on the same inputs (one GPU, run under gpurun):

  configs[0]  "PointNet++ SA/FP ops on one synthetic 20k-point scene, B=1, npoint=2048, nsample=64, CPU reference":
              every operator of the drop-in surface alone, device time by CUDA graph replay (10 launches per replay),
              CPU time of the oracle's C/OpenMP port (all host threads) on the same arrays, results compared.
  configs[2]  "size-adaptive clustering head: 256 vote clusters with per-cluster radius from predicted box size,
              SUN RGB-D shape (20k pts)": votes -> FPS -> size head -> radius -> adaptive ball query -> group -> MLP -> max,
              B = 1 and B = 8 scenes, GPU (graph replay) vs the oracle.

    python tools/config_table.py --out profiles/r02_configs.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200 as S  # noqa: E402
from sad_b200 import mlp as M, ops  # noqa: E402
from sad_b200.config import LAYER_CFG, make_params  # noqa: E402
from sad_b200.modules import SADHotPath  # noqa: E402
from sad_b200.scenes import make_scenes, make_sizes  # noqa: E402
from oracle import c_port as C  # noqa: E402  (tools may time the oracle as the CPU baseline)
from oracle import sad_oracle as O  # noqa: E402

DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def gpu_us(fn, reps=10, it=15):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
        g.replay()
        st.synchronize()
        ts = []
        for _ in range(it):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            g.replay()
            b.record(st)
            st.synchronize()
            ts.append(a.elapsed_time(b) / reps)
    ts.sort()
    return 1e3 * ts[len(ts) // 2]


def cpu_us(fn, reps=3):
    best = 1e30
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return 1e6 * best, out


def config0():
    """One 20k-point scene, npoint 2048, nsample 64 (SA1 of the SUN RGB-D shape) and the FP shapes that follow."""
    rows = []
    xyz_np, feat_np = make_scenes(1, 20000, "surface")
    xyz, feat = cu(xyz_np), cu(feat_np)

    def rec(op, shape, g_us, c_us, exact):
        rows.append({"op": op, "shape": shape, "gpu_us": round(g_us, 1), "cpu_us": round(c_us, 1),
                     "speedup": round(c_us / g_us, 1), "match": exact})

    # a1 furthest point sampling (plain kernel and the scene-grid kernel)
    c_us, inds_np = cpu_us(lambda: C.furthest_point_sample(xyz_np, 2048))
    inds = ops.furthest_point_sample(xyz, 2048)
    rec("furthest_point_sample (cluster kernel)", "1x20000 -> 2048", gpu_us(lambda: ops.furthest_point_sample(xyz, 2048), reps=3),
        c_us, "bit-exact" if np.array_equal(inds.cpu().numpy(), inds_np) else "MISMATCH")
    grid = ops.build_scene_grid(xyz)
    inds_g = ops.furthest_point_sample(xyz, 2048, grid)
    rec("furthest_point_sample (scene grid, culled)", "1x20000 -> 2048",
        gpu_us(lambda: ops.furthest_point_sample(xyz, 2048, grid), reps=3), c_us,
        "bit-exact" if np.array_equal(inds_g.cpu().numpy(), inds_np) else "MISMATCH")
    rec("scene_grid_build", "1x20000", gpu_us(lambda: ops.build_scene_grid(xyz)), float("nan"), "n/a (GPU-side index)")
    # a2 gather
    new_xyz_np = np.stack([xyz_np[0][inds_np[0]]])
    new_xyz = cu(new_xyz_np)
    f64 = np.random.default_rng(0).standard_normal((1, 64, 20000)).astype(np.float32)
    f64g = cu(f64)
    c_us, want = cpu_us(lambda: C.gather_operation(f64, inds_np))
    got = S.gather_operation(f64g, inds)
    rec("gather_operation", "C=64, 20000 -> 2048", gpu_us(lambda: S.gather_operation(f64g, inds)), c_us,
        "bit-exact" if np.array_equal(got.cpu().numpy(), want) else "MISMATCH")
    # a3 ball query (brute force and grid)
    c_us, idx_np = cpu_us(lambda: C.ball_query(0.2, 64, xyz_np, new_xyz_np))
    idx = ops.ball_query(0.2, 64, xyz, new_xyz)
    rec("ball_query (TMA-tiled brute force)", "2048 queries x 20000, r=0.2, nsample 64",
        gpu_us(lambda: ops.ball_query(0.2, 64, xyz, new_xyz)), c_us,
        "bit-exact" if np.array_equal(idx.cpu().numpy(), idx_np) else "MISMATCH")
    idx2 = ops.ball_query(0.2, 64, xyz, new_xyz, grid)
    rec("ball_query (scene grid)", "2048 queries x 20000, r=0.2, nsample 64",
        gpu_us(lambda: ops.ball_query(0.2, 64, xyz, new_xyz, grid)), c_us,
        "bit-exact" if np.array_equal(idx2.cpu().numpy(), idx_np) else "MISMATCH")
    # a4 adaptive-radius ball query
    rad_np = (0.1 + 0.3 * np.random.default_rng(1).random((1, 2048))).astype(np.float32)
    rad = cu(rad_np)
    c_us, idxa_np = cpu_us(lambda: C.ball_query_adaptive(rad_np, 64, xyz_np, new_xyz_np))
    idxa = ops.ball_query_adaptive(rad, 64, xyz, new_xyz)
    rec("ball_query_adaptive", "2048 queries x 20000, r in [0.1,0.4], nsample 64",
        gpu_us(lambda: ops.ball_query_adaptive(rad, 64, xyz, new_xyz)), c_us,
        "bit-exact" if np.array_equal(idxa.cpu().numpy(), idxa_np) else "MISMATCH")
    # a5 grouping
    c_us, want = cpu_us(lambda: C.grouping_operation(f64, idx_np))
    got = S.grouping_operation(f64g, idx)
    rec("grouping_operation", "C=64, (2048,64) of 20000", gpu_us(lambda: S.grouping_operation(f64g, idx)), c_us,
        "bit-exact" if np.array_equal(got.cpu().numpy(), want) else "MISMATCH")
    # a6 fused SA1 stage vs oracle group + MLP + max (NumPy/BLAS MLP)
    params = make_params(0)
    mlp = M.prepare_layers([(cu(W), cu(b)) for W, b in params["sa1"]])

    def cpu_sa():
        x = O.query_and_group(xyz_np, new_xyz_np, feat_np, idx_np, np.float32(0.2), True, True, impl=C)
        return O.shared_mlp(x, params["sa1"], pool=True)
    c_us, want = cpu_us(cpu_sa, reps=2)
    got = M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.2, mlp).cpu().numpy()
    err = float(np.abs(got - want).max() / np.abs(want).max())
    rec("group + shared MLP [4,64,64,128] + max-pool (fused, bf16)", "2048 x 64 rows",
        gpu_us(lambda: M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.2, mlp)), c_us, f"rel err {err:.2e} (bar 2e-2)")
    mlp32 = M.prepare_layers([(cu(W), cu(b)) for W, b in params["sa1"]], dtype="tf32")
    got = M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.2, mlp32).cpu().numpy()
    err = float(np.abs(got - want).max() / np.abs(want).max())
    rec("group + shared MLP [4,64,64,128] + max-pool (fused, tf32)", "2048 x 64 rows",
        gpu_us(lambda: M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.2, mlp32)), c_us, f"rel err {err:.2e} (bar 2e-3)")
    # a8 / a9 feature propagation: 2048 unknown <- 512 known, C = 256
    known_np = new_xyz_np[:, :512].copy()
    known = cu(known_np)
    c_us, (d_np, i_np) = cpu_us(lambda: C.three_nn(new_xyz_np, known_np))
    d, i = S.three_nn(new_xyz, known)
    ok = np.array_equal(i.cpu().numpy(), i_np) and np.array_equal(d.cpu().numpy(), d_np)
    rec("three_nn", "2048 unknown x 512 known", gpu_us(lambda: S.three_nn(new_xyz, known)), c_us,
        "bit-exact (idx and dist)" if ok else "MISMATCH")
    w_np = O.interpolation_weights(d_np)
    kf = np.random.default_rng(2).standard_normal((1, 256, 512)).astype(np.float32)
    kfg, wg, ig = cu(kf), cu(w_np), cu(i_np.astype(np.int32))
    c_us, want = cpu_us(lambda: C.three_interpolate(kf, i_np, w_np))
    got = S.three_interpolate(kfg, ig, wg)
    rec("three_interpolate", "C=256, 512 -> 2048", gpu_us(lambda: S.three_interpolate(kfg, ig, wg)), c_us,
        "bit-exact" if np.array_equal(got.cpu().numpy(), want) else "MISMATCH")
    return rows


def config2():
    """Clustering head alone on the votes of 20k-point scenes, with the size head predicting the box sizes."""
    rows = []
    params = make_params(0)
    cfg = LAYER_CFG
    model = SADHotPath(1).load_params(params).to(DEV).eval()
    for B in (1, 8):
        xyz_np, feat_np = make_scenes(B, 20000, "surface")
        with torch.no_grad():
            end = model(cu(xyz_np), cu(feat_np))
            vxyz, vfeat = end["vote_xyz"].contiguous(), end["vote_features"].contiguous()

            def head():
                return model.agg(vxyz, vfeat, None)
            g_us = gpu_us(head, reps=4)
            cxyz, cfeat, cinds, radius_t, size = head()
        vx, vf = vxyz.cpu().numpy(), vfeat.cpu().numpy()
        npoint, _, nsample = cfg["agg"]

        def cpu_head():
            ci = C.furthest_point_sample(vx, npoint)
            centre = np.stack([vf[b][:, ci[b]] for b in range(B)])
            sz = O.size_head(centre, params["size"], cfg["size_scale"], cfg["size_clip"])
            return O.vote_aggregation(vx, vf, sz, npoint, nsample, params["agg"], alpha=cfg["alpha"], r_min=cfg["r_min"],
                                      r_max=cfg["r_max"], impl=C), sz
        c_us, ((wxyz, wfeat, winds, wrad), wsize) = cpu_us(cpu_head, reps=2)
        # the size head runs through the bf16 MLP: compare the rest on the GPU's own sizes
        oxyz, ofeat, oinds, orad = O.vote_aggregation(vx, vf, size.cpu().numpy(), npoint, nsample, params["agg"], alpha=cfg["alpha"],
                                                      r_min=cfg["r_min"], r_max=cfg["r_max"], impl=C)
        err = float(np.abs(cfeat.cpu().numpy() - ofeat).max() / np.abs(ofeat).max())
        size_err = float(np.abs(size.cpu().numpy() - wsize).max() / np.abs(wsize).max())
        rows.append({"scenes": B, "points_per_scene": 20000, "votes": int(vx.shape[1]), "clusters": npoint, "nsample": nsample,
                     "gpu_us": round(g_us, 1), "gpu_scenes_per_s": round(B / (g_us * 1e-6)),
                     "cpu_us": round(c_us, 1), "cpu_scenes_per_s": round(B / (c_us * 1e-6), 1),
                     "cluster_inds": "bit-exact" if np.array_equal(cinds.cpu().numpy(), oinds) else "MISMATCH",
                     "radius": "bit-exact" if np.array_equal(radius_t.cpu().numpy(), orad) else "MISMATCH",
                     "cluster_features_rel_err": round(err, 5), "predicted_size_rel_err": round(size_err, 5),
                     "launches": "FPS(votes) + gather + size head MLP + size_to_radius + gather_points + "
                                 "ball_query_adaptive + fused group/MLP/max"})
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    a = ap.parse_args()
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count()
    out = {"gpu": torch.cuda.get_device_name(0), "host_threads": threads,
           "timing": "GPU: CUDA graph of N launches replayed, median of 15, per launch; CPU: best of 2-3 passes of the oracle's "
                     "C/OpenMP port (NumPy/BLAS for the MLP) on all host threads",
           "configs[0] one 20k-point scene, per operator": config0(),
           "configs[2] size-adaptive clustering head": config2()}
    json.dump(out, open(a.out, "w"), indent=1)
    for k in ("configs[0] one 20k-point scene, per operator", "configs[2] size-adaptive clustering head"):
        print(k)
        for r in out[k]:
            print("  ", r)


if __name__ == "__main__":
    main()
