"""Generates tests/golden/hot_path_small.npz -- seeded inputs and the oracle's outputs for every op of the
hot path on a small scene.

The mounted reference has no code, tests or fixtures (README.md:1-2 only), so these vectors cannot come from it:
they are produced by oracle/sad_oracle.py (NumPy, written from the ops' definitions) and pin BOTH the oracle
(tests/test_golden.py, CPU: NumPy file and C port must reproduce them bit for bit, so an accidental change of the
checker is caught) and the CUDA kernels (tests/test_golden.py, -m gpu).

    python tests/golden/make_golden.py        # rewrites the fixture; commit the result
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sad_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hot_path_small.npz")


def inputs():
    rng = np.random.default_rng(20261018)
    B, N = 2, 700
    # a lattice-quantised half (exact ties, duplicates) and a continuous half
    xyz = (rng.random((B, N, 3), dtype=np.float32) * np.array([4, 4, 2], np.float32)).astype(np.float32)
    xyz[:, : N // 2] = np.round(xyz[:, : N // 2] / np.float32(0.25)) * np.float32(0.25)
    feat = rng.standard_normal((B, 6, N)).astype(np.float32)
    size = rng.uniform(0.2, 2.0, (B, 24, 3)).astype(np.float32)
    layers_sa = [((rng.standard_normal((co, ci)) / np.sqrt(ci)).astype(np.float32),
                  (rng.standard_normal(co) * 0.1).astype(np.float32)) for ci, co in [(9, 64), (64, 64), (64, 32)]]
    return xyz, feat, size, layers_sa


def main():
    xyz, feat, size, layers = inputs()
    d = {"xyz": xyz, "feat": feat, "size": size}
    for i, (W, b) in enumerate(layers):
        d[f"W{i}"], d[f"b{i}"] = W, b
    d["fps_idx"] = O.furthest_point_sample(xyz, 96)
    new_xyz = np.stack([xyz[b][d["fps_idx"][b]] for b in range(xyz.shape[0])])
    d["new_xyz"] = new_xyz
    d["gather"] = O.gather_operation(feat, d["fps_idx"])
    d["bq_idx"] = O.ball_query(0.45, 16, xyz, new_xyz)
    rt = O.size_to_radius(size[:, :24], 1.0, 0.1, 1.2)
    d["radius_t"] = rt
    d["bqa_idx"] = O.ball_query_adaptive(rt, 8, xyz, new_xyz[:, :24])
    d["grouped"] = O.grouping_operation(feat, d["bq_idx"])
    dist, nn = O.three_nn(xyz, new_xyz)
    d["nn_dist"], d["nn_idx"] = dist, nn
    w = O.interpolation_weights(dist)
    d["nn_weight"] = w
    d["interp"] = O.three_interpolate(d["gather"], nn, w)
    x = O.query_and_group(xyz, new_xyz, feat, d["bq_idx"], np.float32(0.45), True, True)
    d["sa_features_f32"] = O.shared_mlp(x, layers, pool=True)
    d["sa_features_bf16"] = O.shared_mlp(x, layers, pool=True, emulate_bf16=True)
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, {k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    main()
