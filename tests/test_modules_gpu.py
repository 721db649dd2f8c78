"""Module / backbone parity (tier "module parity"): SA, FP, the whole
backbone and the size-adaptive clustering head vs the same modules built from the oracle's
ops, same weights.  Every index tensor that depends only on coordinates must be bit-exact;
features go through the bf16 MLP and are held to BASELINE's 2e-2 bar (relative to the
feature scale of the layer)."""
import numpy as np
import pytest
import torch

from oracle import sad_oracle as O
from oracle import c_port as C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def close(got, want, tol=2e-2):
    got = got.detach().float().cpu().numpy()
    scale = max(1e-6, float(np.abs(want).max()))
    err = float(np.abs(got - want).max())
    assert err <= tol * scale, f"max abs err {err:.4g} vs scale {scale:.4g} (tol {tol})"


@pytest.fixture(scope="module")
def setup():
    import sad_b200  # noqa: F401
    from sad_b200.config import LAYER_CFG, make_params
    from sad_b200.modules import SADHotPath
    from sad_b200.scenes import make_scenes, make_sizes
    params = make_params(0)
    model = SADHotPath(1).load_params(params).to(DEV).eval()
    return model, params, LAYER_CFG, make_scenes, make_sizes


@pytest.mark.parametrize("N,kind", [(5000, "surface"), (20000, "uniform")])
def test_hot_path_matches_oracle(setup, N, kind):
    model, params, cfg, make_scenes, make_sizes = setup
    B = 2
    xyz, feat = make_scenes(B, N, kind)
    size = make_sizes(B, cfg["agg"][0])
    want = O.backbone_forward(xyz, feat, params, cfg, impl=C)
    with torch.no_grad():
        got = model(cu(xyz), cu(feat), cu(size))
    for name in ("sa1", "sa2", "sa3", "sa4"):
        assert np.array_equal(got[name + "_inds"].cpu().numpy(), want[name + "_inds"]), name
        assert np.array_equal(got[name + "_xyz"].cpu().numpy(), want[name + "_xyz"]), name
        close(got[name + "_features"], want[name + "_features"])
    close(got["fp2_features"], want["fp2_features"])
    assert np.array_equal(got["fp2_inds"].cpu().numpy(), want["fp2_inds"])

    # clustering head: feed the GPU's own votes to the oracle so that indices must match exactly
    vxyz = got["vote_xyz"].cpu().numpy()
    vfeat = got["vote_features"].cpu().numpy()
    wv_xyz, wv_feat = O.voting_module(want["fp2_xyz"], got["fp2_features"].cpu().numpy(), params["vote"])
    close(got["vote_xyz"], wv_xyz)
    close(got["vote_features"], wv_feat)
    npoint, _, nsample = cfg["agg"]
    cxyz, cfeat, cinds, rt = O.vote_aggregation(vxyz, vfeat, size, npoint, nsample, params["agg"],
                                                alpha=cfg["alpha"], r_min=cfg["r_min"], r_max=cfg["r_max"], impl=C)
    assert np.array_equal(got["cluster_radius"].cpu().numpy(), rt)
    assert np.array_equal(got["cluster_inds"].cpu().numpy(), cinds)
    assert np.array_equal(got["cluster_xyz"].cpu().numpy(), cxyz)
    close(got["cluster_features"], cfeat)


def test_train_mode_matches_eval_and_backprops(setup):
    """The torch (training) composition and the fused eval path agree; gradients flow through
    the scatter-add backward kernels to the input features."""
    from sad_b200.modules import PointnetSAModuleVotes, PointnetFPModule
    rng = np.random.default_rng(3)
    xyz = cu((rng.random((2, 600, 3), dtype=np.float32) * 2).astype(np.float32))
    feat = cu(rng.standard_normal((2, 6, 600)).astype(np.float32)).requires_grad_(True)
    sa = PointnetSAModuleVotes(64, 0.5, 8, [6, 64, 32]).to(DEV)
    fp = PointnetFPModule([32 + 6, 16]).to(DEV)
    sa.eval(), fp.eval()
    new_xyz, f_train_path, inds = sa(xyz, feat)                    # grad enabled -> torch path
    up = fp(xyz, new_xyz, feat, f_train_path)
    up.square().mean().backward()
    assert feat.grad is not None and torch.isfinite(feat.grad).all() and feat.grad.abs().sum() > 0
    with torch.no_grad():
        _, f_eval_path, inds2 = sa(xyz, feat.detach())
    assert torch.equal(inds, inds2)
    close(f_eval_path, f_train_path.detach().cpu().numpy())


def test_forward_host_roundtrip(setup):
    model, params, cfg, make_scenes, make_sizes = setup
    xyz, feat = make_scenes(1, 4096, "surface")
    size = make_sizes(1, cfg["agg"][0])
    host = [torch.from_numpy(a).pin_memory() for a in (xyz, feat, size)]
    cx, cf = model.forward_host(*host)
    with torch.no_grad():
        end = model(cu(xyz), cu(feat), cu(size))
    assert torch.equal(cx, end["cluster_xyz"].cpu()) and torch.equal(cf, end["cluster_features"].cpu())
    assert not cx.is_cuda and tuple(cf.shape) == (1, 128, 256)


def test_folded_weights_follow_parameter_updates():
    """the BN-folded / packed weight cache must not survive load_state_dict or an in-place update."""
    from sad_b200.modules import SharedMLP
    from sad_b200 import mlp as M
    torch.manual_seed(0)
    sm = SharedMLP([64, 64, 64]).to(DEV).eval()
    x = torch.randn(2, 64, 128, device=DEV)
    with torch.no_grad():
        y0 = M.pointwise_mlp(x, sm.folded()).clone()
        sd = {k: (v * 0.5 if k.startswith("convs") else v) for k, v in sm.state_dict().items()}
        sm.load_state_dict(sd)
        y1 = M.pointwise_mlp(x, sm.folded()).clone()
        ref = sm(x.unsqueeze(-1)).squeeze(-1)
        assert not torch.allclose(y0, y1), "eval forward after load_state_dict still ran the old weights"
        assert float((y1 - ref).abs().max()) <= 2e-2 * max(1e-6, float(ref.abs().max()))
        sm.convs[0].weight.mul_(2.0)                   # in-place update (an optimizer step)
        y2 = M.pointwise_mlp(x, sm.folded())
        ref2 = sm(x.unsqueeze(-1)).squeeze(-1)
        assert float((y2 - ref2).abs().max()) <= 2e-2 * max(1e-6, float(ref2.abs().max()))


def test_size_head_drives_the_adaptive_radius():
    """SURVEY 8(f) rank 3 / VERDICT r1 missing item 4: with no sizes passed in, the aggregation module predicts them
    from the vote features at the cluster centres (size head through the fused kernel) and feeds the radius of the
    adaptive ball query.  Predicted sizes within the bf16 bar of the oracle's size head; with the GPU's own sizes the
    oracle's clustering indices match bit for bit."""
    import sad_b200  # noqa: F401
    from sad_b200.config import LAYER_CFG, make_params
    from sad_b200.modules import SADHotPath
    from sad_b200.scenes import make_scenes
    params = make_params(0)
    model = SADHotPath(1).load_params(params).to(DEV).eval()
    xyz, feat = make_scenes(2, 9000, "surface", first_scene=5)
    with torch.no_grad():
        got = model(torch.from_numpy(xyz).to(DEV), torch.from_numpy(feat).to(DEV))          # size=None: the head
    vfeat = got["vote_features"].cpu().numpy()
    vxyz = got["vote_xyz"].cpu().numpy()
    cinds = got["cluster_inds"].cpu().numpy()
    centre = np.stack([vfeat[b][:, cinds[b]] for b in range(2)])
    want_size = O.size_head(centre, params["size"], LAYER_CFG["size_scale"], LAYER_CFG["size_clip"])
    size = got["cluster_size"].cpu().numpy()
    assert size.shape == (2, LAYER_CFG["agg"][0], 3)
    assert float(np.abs(size - want_size).max()) <= 3e-2 * float(np.abs(want_size).max())
    assert float(size.std()) > 1e-3                                   # the head actually differentiates clusters
    npoint, _, nsample = LAYER_CFG["agg"]
    cxyz, cfeat, winds, rt = O.vote_aggregation(vxyz, vfeat, size, npoint, nsample, params["agg"], alpha=LAYER_CFG["alpha"],
                                                r_min=LAYER_CFG["r_min"], r_max=LAYER_CFG["r_max"], impl=C)
    assert np.array_equal(cinds, winds) and np.array_equal(got["cluster_radius"].cpu().numpy(), rt)
    assert np.array_equal(got["cluster_xyz"].cpu().numpy(), cxyz)
    close(got["cluster_features"], cfeat)
    assert float(rt.min()) < float(rt.max())                          # radii spread inside the clamp range


def test_channel_last_twin_is_dropped_after_an_in_place_edit():
    """a fused stage leaves a bf16 channel-last twin on its f32 output for the next stage; editing the f32
    tensor in place must invalidate it."""
    from sad_b200 import mlp as M
    torch.manual_seed(1)
    layers = [(torch.randn(64, 64, device=DEV) / 8, torch.zeros(64, device=DEV)) for _ in range(2)]
    mlp = M.prepare_layers(layers)
    x = torch.randn(2, 64, 256, device=DEV)
    y = M.pointwise_mlp(x, mlp)
    assert M.to_cl_bf16(y) is y._sad_cl
    z0 = M.pointwise_mlp(y, mlp).clone()
    y.mul_(2.0)
    assert M.to_cl_bf16(y) is not y._sad_cl
    z1 = M.pointwise_mlp(y, mlp)
    assert not torch.allclose(z0, z1)
