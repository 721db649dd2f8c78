"""Golden vectors (tests/golden/hot_path_small.npz, made by tests/golden/make_golden.py from the NumPy
oracle -- the reference mount has no fixtures to use instead).  CPU: both oracle implementations reproduce
them bit for bit.  GPU (-m gpu): the CUDA kernels do -- indices and fp32 copies exactly, the bf16 MLP
within BASELINE's 2e-2."""
import os

import numpy as np
import pytest

from oracle import sad_oracle as O
from oracle import c_port as C

G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hot_path_small.npz")))
LAYERS = [(G[f"W{i}"], G[f"b{i}"]) for i in range(3)]


@pytest.mark.parametrize("impl", [O, C], ids=["numpy", "c_port"])
def test_oracle_reproduces_golden(impl):
    xyz, feat, new_xyz = G["xyz"], G["feat"], G["new_xyz"]
    assert np.array_equal(impl.furthest_point_sample(xyz, 96), G["fps_idx"])
    assert np.array_equal(impl.ball_query(0.45, 16, xyz, new_xyz), G["bq_idx"])
    assert np.array_equal(impl.ball_query_adaptive(G["radius_t"], 8, xyz, new_xyz[:, :24]), G["bqa_idx"])
    assert np.array_equal(impl.grouping_operation(feat, G["bq_idx"]), G["grouped"])
    assert np.array_equal(impl.gather_operation(feat, G["fps_idx"]), G["gather"])
    dist, nn = impl.three_nn(xyz, new_xyz)
    assert np.array_equal(nn, G["nn_idx"]) and np.array_equal(dist, G["nn_dist"])
    assert np.array_equal(impl.three_interpolate(G["gather"], G["nn_idx"], G["nn_weight"]), G["interp"])


def test_oracle_mlp_and_radius_reproduce_golden():
    assert np.array_equal(O.size_to_radius(G["size"], 1.0, 0.1, 1.2), G["radius_t"])
    x = O.query_and_group(G["xyz"], G["new_xyz"], G["feat"], G["bq_idx"], np.float32(0.45), True, True)
    np.testing.assert_allclose(O.shared_mlp(x, LAYERS, pool=True), G["sa_features_f32"], rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_cuda_kernels_reproduce_golden():
    import torch
    import sad_b200 as S
    from sad_b200 import mlp as M, ops
    dev = "cuda:0"
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)      # noqa: E731
    xyz, feat, new_xyz = cu(G["xyz"]), cu(G["feat"]), cu(G["new_xyz"])
    fps = S.furthest_point_sample(xyz, 96)
    assert np.array_equal(fps.cpu().numpy(), G["fps_idx"])
    grid = ops.build_scene_grid(xyz)                                      # the culled / grid kernels too
    assert np.array_equal(S.furthest_point_sample(xyz, 96, grid).cpu().numpy(), G["fps_idx"])
    assert np.array_equal(S.gather_operation(feat, fps).cpu().numpy(), G["gather"])
    for g in (None, grid):
        assert np.array_equal(S.ball_query(0.45, 16, xyz, new_xyz, g).cpu().numpy(), G["bq_idx"])
        assert np.array_equal(S.ball_query_adaptive(cu(G["radius_t"]), 8, xyz, new_xyz[:, :24].contiguous(), g).cpu().numpy(),
                              G["bqa_idx"])
    idx = cu(G["bq_idx"])
    assert np.array_equal(S.grouping_operation(feat, idx).cpu().numpy(), G["grouped"])
    dist, nn = S.three_nn(xyz, new_xyz)
    assert np.array_equal(nn.cpu().numpy(), G["nn_idx"]) and np.array_equal(dist.cpu().numpy(), G["nn_dist"])
    got = S.three_interpolate(cu(G["gather"]), nn, cu(G["nn_weight"]))
    assert np.array_equal(got.cpu().numpy(), G["interp"])
    assert np.array_equal(S.size_to_radius(cu(G["size"]), 1.0, 0.1, 1.2).cpu().numpy(), G["radius_t"])
    mlp = M.prepare_layers([(cu(W), cu(b)) for W, b in LAYERS])
    f = M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.45, mlp, use_xyz=True, normalize_xyz=True).cpu().numpy()
    scale = float(np.abs(G["sa_features_f32"]).max())
    assert float(np.abs(f - G["sa_features_f32"]).max()) <= 2e-2 * scale
    assert float(np.abs(f - G["sa_features_bf16"]).max()) <= 8e-3 * scale
