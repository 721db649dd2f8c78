"""a6 in tf32 mode (csrc/mlp_tf32.cu, `PreparedMLP(dtype="tf32")`): the fused gather / interpolation -> MLP -> max-pool
stage with fp32 channel-last activations and kind::tf32 MMAs, against the fp32 oracle.

Bar: tf32 keeps a 10-bit mantissa (operands rounded to nearest, 2^-11 relative each), accumulation is fp32; over K <= 512
products and three layers the error stays below 2e-3 of the layer's feature scale (norm-wise, `close`), ten times
tighter than the bf16 path's 2e-2.  Indices never depend on the MLP precision."""
import numpy as np
import pytest
import torch

from oracle import sad_oracle as O
from oracle import c_port as C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 2e-3


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def close(got, want, tol=TOL):
    got = got.detach().float().cpu().numpy()
    scale = max(1e-6, float(np.abs(want).max()))
    err = float(np.abs(got - want).max())
    assert err <= tol * scale, f"max abs err {err:.4g} vs scale {scale:.4g} (tol {tol})"


def make_layers(rng, chans, bias_std=0.1):
    return [((rng.standard_normal((co, ci)) / np.sqrt(ci)).astype(np.float32),
             (bias_std * rng.standard_normal(co)).astype(np.float32)) for ci, co in zip(chans[:-1], chans[1:])]


def tlayers(layers):
    return [(cu(W), cu(b)) for W, b in layers]


@pytest.mark.parametrize("B,N,P,S,Cf,hidden,adaptive,use_xyz", [
    (2, 3000, 100, 16, 64, [64, 64, 128], False, True),        # partial last tile
    (1, 4000, 128, 64, 1, [64, 64, 128], False, True),         # SA1 shape: one scalar feature in the special K step
    (2, 2048, 96, 32, 128, [128, 128, 256], False, True),      # SA2 shape
    (2, 1024, 64, 16, 256, [128, 128, 256], False, True),      # SA3 / SA4 shape, 9 K chunks
    (2, 1024, 50, 16, 256, [128, 128, 128], True, True),       # aggregation: per-cluster radius
    (1, 2000, 77, 8, 36, [32, 96], False, True),               # two layers, odd widths (32-multiples / last 96)
    (1, 2000, 64, 16, 64, [64, 64, 40], False, False),         # no xyz columns
    (1, 1500, 40, 128, 0, [32, 32, 64], False, True),          # coordinates only, nsample 128
    (3, 2500, 256, 16, 128, [256, 256, 512], False, True),     # widest layers: 4 last-layer blocks share the hidden columns
])
def test_sa_stage_tf32(B, N, P, S, Cf, hidden, adaptive, use_xyz):
    from sad_b200 import mlp as M
    rng = np.random.default_rng(B * 1000 + P + S + Cf)
    xyz = (rng.random((B, N, 3), dtype=np.float32) * 3).astype(np.float32)
    feat = rng.standard_normal((B, Cf, N)).astype(np.float32) if Cf else None
    inds = C.furthest_point_sample(xyz, P)
    new_xyz = np.stack([xyz[b][inds[b]] for b in range(B)])
    if adaptive:
        rad = (0.3 + 0.5 * rng.random((B, P))).astype(np.float32)
        idx = C.ball_query_adaptive(rad, S, xyz, new_xyz)
        rad_o, rad_g = rad, cu(rad)
    else:
        idx = C.ball_query(0.5, S, xyz, new_xyz)
        rad_o, rad_g = np.float32(0.5), 0.5
    layers = make_layers(rng, [Cf + (3 if use_xyz or not Cf else 0)] + hidden)
    x = O.query_and_group(xyz, new_xyz, feat, idx, rad_o, use_xyz, True)
    want = O.shared_mlp(x, layers, pool=True)
    mlp = M.prepare_layers(tlayers(layers), dtype="tf32")
    assert mlp.fusable(S)
    got = M.sa_group_mlp(cu(xyz), cu(new_xyz), None if feat is None else cu(feat), cu(idx), rad_g, mlp, use_xyz=use_xyz,
                         normalize_xyz=True)
    assert tuple(got.shape) == (B, hidden[-1], P) and got.dtype == torch.float32
    close(got, want)
    twin = got._sad_cl32                      # fp32 channel-last twin: the next tf32 stage's source
    assert tuple(twin.shape) == (B, P, hidden[-1]) and twin.dtype == torch.float32
    assert torch.equal(twin.transpose(1, 2), got)


@pytest.mark.parametrize("B,n,m,C2,C1,hidden", [
    (2, 200, 64, 256, 128, [256, 256]),       # FP shape, n not a multiple of 128
    (2, 512, 256, 256, 256, [256, 256]),      # FP1
    (1, 130, 40, 64, 0, [64, 32, 48]),        # no skip features, three layers
])
def test_fp_stage_tf32_interpolates_in_the_kernel(B, n, m, C2, C1, hidden):
    from sad_b200 import mlp as M
    import sad_b200 as S_
    rng = np.random.default_rng(n + m)
    unknown = (rng.random((B, n, 3), dtype=np.float32) * 2).astype(np.float32)
    known = (rng.random((B, m, 3), dtype=np.float32) * 2).astype(np.float32)
    kf = rng.standard_normal((B, C2, m)).astype(np.float32)
    sf = rng.standard_normal((B, C1, n)).astype(np.float32) if C1 else None
    dist, idx = O.three_nn(unknown, known)
    w = O.interpolation_weights(dist)
    interp = O.three_interpolate(kf, idx, w)
    x = interp if sf is None else np.concatenate([interp, sf], axis=1)
    layers = make_layers(rng, [C2 + C1] + hidden)
    want = O.shared_mlp(x[..., None], layers, pool=False)[..., 0]
    mlp = M.prepare_layers(tlayers(layers), dtype="tf32")
    got = M.fp_interp_mlp(cu(kf), None if sf is None else cu(sf), cu(idx.astype(np.int32)), cu(w), mlp)
    assert tuple(got.shape) == (B, hidden[-1], n)
    close(got, want)
    assert S_ is not None


def test_pointwise_tf32_linear_last_layer_three_blocks():
    """voting MLP shape: 256 -> 256 -> 256 -> 3 + 256, last layer linear (259 channels = three 128-channel blocks)."""
    from sad_b200 import mlp as M
    rng = np.random.default_rng(5)
    B, n = 2, 300
    x = rng.standard_normal((B, 256, n)).astype(np.float32)
    layers = make_layers(rng, [256, 256, 256, 259])
    want = O.shared_mlp(x[..., None], layers, pool=False, last_relu=False)[..., 0]
    mlp = M.prepare_layers(tlayers(layers), dtype="tf32")
    got = M.pointwise_mlp(cu(x), mlp, last_relu=False)
    close(got, want)
    assert float(got.min()) < 0          # linear: negative outputs survive


def test_tf32_is_tighter_than_bf16_on_the_same_stage():
    from sad_b200 import mlp as M
    rng = np.random.default_rng(11)
    B, N, P, S = 2, 2048, 128, 32
    xyz = (rng.random((B, N, 3), dtype=np.float32) * 3).astype(np.float32)
    feat = rng.standard_normal((B, 128, N)).astype(np.float32)
    inds = C.furthest_point_sample(xyz, P)
    new_xyz = np.stack([xyz[b][inds[b]] for b in range(B)])
    idx = C.ball_query(0.5, S, xyz, new_xyz)
    layers = make_layers(rng, [131, 128, 128, 256])
    want = O.shared_mlp(O.query_and_group(xyz, new_xyz, feat, idx, np.float32(0.5), True, True), layers, pool=True)
    errs = {}
    for dt in ("bf16", "tf32"):
        mlp = M.prepare_layers(tlayers(layers), dtype=dt)
        got = M.sa_group_mlp(cu(xyz), cu(new_xyz), cu(feat), cu(idx), 0.5, mlp).cpu().numpy()
        errs[dt] = float(np.abs(got - want).max() / np.abs(want).max())
    assert errs["tf32"] < TOL and errs["tf32"] < 0.25 * errs["bf16"], errs


def test_hot_path_in_tf32_mode_matches_oracle():
    """The whole path with every MLP stage in tf32 mode: indices bit-exact, features at the tf32 bar stage by stage
    (each stage compared on the oracle's continuation of the GPU's own inputs would hide nothing here: the error of
    eight chained stages stays under 5e-3 of the feature scale)."""
    import sad_b200  # noqa: F401
    from sad_b200.config import LAYER_CFG as cfg, make_params
    from sad_b200.modules import SADHotPath
    from sad_b200.scenes import make_scenes, make_sizes
    params = make_params(0)
    model = SADHotPath(1, mlp_dtype="tf32").load_params(params).to(DEV).eval()
    B, N = 2, 6000
    xyz, feat = make_scenes(B, N, "surface")
    size = make_sizes(B, cfg["agg"][0])
    want = O.backbone_forward(xyz, feat, params, cfg, impl=C)
    with torch.no_grad():
        got = model(cu(xyz), cu(feat), cu(size))
    for name in ("sa1", "sa2", "sa3", "sa4"):
        assert np.array_equal(got[name + "_inds"].cpu().numpy(), want[name + "_inds"]), name
        close(got[name + "_features"], want[name + "_features"], 5e-3)
    close(got["fp2_features"], want["fp2_features"], 5e-3)
    vxyz = got["vote_xyz"].cpu().numpy()
    vfeat = got["vote_features"].cpu().numpy()
    wv_xyz, wv_feat = O.voting_module(want["fp2_xyz"], got["fp2_features"].cpu().numpy(), params["vote"])
    close(got["vote_xyz"], wv_xyz)
    close(got["vote_features"], wv_feat)
    npoint, _, nsample = cfg["agg"]
    cxyz, cfeat, cinds, rt = O.vote_aggregation(vxyz, vfeat, size, npoint, nsample, params["agg"],
                                                alpha=cfg["alpha"], r_min=cfg["r_min"], r_max=cfg["r_max"], impl=C)
    assert np.array_equal(got["cluster_inds"].cpu().numpy(), cinds)
    close(got["cluster_features"], cfeat)
    # and the two precisions agree with each other at the bf16 bar
    model.set_mlp_dtype("bf16")
    with torch.no_grad():
        got16 = model(cu(xyz), cu(feat), cu(size))
    close(got16["fp2_features"], got["fp2_features"].cpu().numpy(), 2e-2)


def test_tf32_entry_point_rejects_bad_arguments():
    import ctypes
    from sad_b200 import _lib
    lib = _lib.load()
    vp = ctypes.c_void_p
    x = torch.zeros(1024, device=DEV)
    st = vp(torch.cuda.current_stream().cuda_stream)
    n = 2
    imgs = (vp * n)(x.data_ptr(), x.data_ptr())
    cout = (ctypes.c_int * n)(48, 64)            # hidden width not a multiple of 32
    rc = lib.sad_mlp_tf32_fwd(1, 128, 128, 1, vp(0), 0, 0, vp(0), vp(0), vp(x.data_ptr()), 32, vp(0), vp(0), vp(0), 0.0, vp(0), 0,
                              vp(0), 0, n, imgs, imgs, cout, 1, vp(x.data_ptr()), vp(0), st)
    assert rc == -3
    cout = (ctypes.c_int * n)(64, 64)
    rc = lib.sad_mlp_tf32_fwd(1, 128, 128, 3, vp(0), 0, 0, vp(0), vp(0), vp(x.data_ptr()), 32, vp(x.data_ptr()), vp(0), vp(0), 0.0,
                              vp(0), 0, vp(0), 0, n, imgs, imgs, cout, 1, vp(x.data_ptr()), vp(0), st)
    assert rc == -3                                # nsample not a power of two
    rc = lib.sad_mlp_tf32_fwd(1, 128, 64, 1, vp(0), 0, 0, vp(0), vp(0), vp(x.data_ptr()), 32, vp(0), vp(0), vp(0), 0.0, vp(0), 0,
                              vp(0), 0, n, imgs, imgs, cout, 1, vp(x.data_ptr()), vp(0), st)
    assert rc == -1                                # identity rows need N == P
