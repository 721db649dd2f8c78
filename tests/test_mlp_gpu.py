"""Fused gather + tcgen05 MLP + max-pool kernel vs the oracle.

Two bars per case: 2e-2 of the layer's feature scale against the fp32 oracle (BASELINE's
bf16 tolerance) and a tighter 8e-3 against the oracle run with bf16-rounded operands
(same storage precision as the kernel, so only the accumulation order differs)."""
import numpy as np
import pytest
import torch

from oracle import sad_oracle as O
from oracle import c_port as C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def close(got, want, tol):
    got = got.detach().float().cpu().numpy()
    scale = max(1e-6, float(np.abs(want).max()))
    err = float(np.abs(got - want).max())
    assert err <= tol * scale, f"max abs err {err:.4g} vs scale {scale:.4g} (tol {tol})"


def make_layers(rng, chans, bias_std=0.1):
    return [((rng.standard_normal((co, ci)) / np.sqrt(ci)).astype(np.float32),
             (rng.standard_normal(co) * bias_std).astype(np.float32)) for ci, co in zip(chans[:-1], chans[1:])]


def tlayers(layers):
    return [(cu(W), cu(b)) for W, b in layers]


@pytest.mark.parametrize("B,N,P,S,Cf,hidden,radius,adaptive", [
    (2, 3000, 64, 64, 1, [64, 64, 128], 0.5, False),        # SA1 shape: fp32 scalar feature in the special chunk
    (2, 2048, 128, 32, 128, [128, 128, 256], 0.6, False),   # SA2
    (1, 1024, 64, 16, 256, [128, 128, 256], 0.9, False),    # SA3 / SA4
    (2, 1024, 96, 16, 256, [128, 128, 128], 0.3, True),     # vote aggregation, per-cluster radius
    (1, 500, 37, 8, 64, [64, 192], 0.8, False),             # 2 layers, partial last tile, odd P
    (1, 300, 5, 2, 5, [64, 64, 70], 1.0, False),            # tiny: c_last not a multiple of anything
    (4, 2048, 1024, 32, 128, [128, 128, 256], 0.6, False),  # SA2 at depth: 7 tiles per CTA, streamed weights, both contexts
    (8, 1024, 512, 16, 256, [128, 128, 256], 0.9, False),   # SA3 at depth: more K chunks than A-ring stages, 3-4 tiles per CTA
    (2, 20000, 2048, 64, 1, [64, 64, 128], 0.3, False),     # SA1 at depth: 28 tiles per CTA, everything pinned
])
def test_fused_sa_stage(B, N, P, S, Cf, hidden, radius, adaptive):
    from sad_b200 import mlp as M
    rng = np.random.default_rng(N + P + S)
    xyz = (rng.random((B, N, 3), dtype=np.float32) * 3).astype(np.float32)
    feat = rng.standard_normal((B, Cf, N)).astype(np.float32)
    inds = C.furthest_point_sample(xyz, P)
    new_xyz = np.stack([xyz[b][inds[b]] for b in range(B)])
    if adaptive:
        rt = (rng.random((B, P), dtype=np.float32) * 0.8 + 0.2).astype(np.float32)
        idx = C.ball_query_adaptive(rt, S, xyz, new_xyz)
        rad_o, rad_g = rt, cu(rt)
    else:
        idx = C.ball_query(radius, S, xyz, new_xyz)
        rad_o, rad_g = np.float32(radius), radius
    layers = make_layers(rng, [Cf + 3] + hidden)
    x = O.query_and_group(xyz, new_xyz, feat, idx, rad_o, True, True)
    want = O.shared_mlp(x, layers, pool=True)
    want_bf = O.shared_mlp(x, layers, pool=True, emulate_bf16=True)
    mlp = M.prepare_layers(tlayers(layers))
    assert mlp.fusable(S)
    got = M.sa_group_mlp(cu(xyz), cu(new_xyz), cu(feat), cu(idx), rad_g, mlp, use_xyz=True, normalize_xyz=True)
    assert tuple(got.shape) == (B, hidden[-1], P) and got.dtype == torch.float32
    close(got, want, 2e-2)
    close(got, want_bf, 8e-3)
    # channel-last bf16 twin carries the same values
    twin = got._sad_cl
    assert tuple(twin.shape) == (B, P, hidden[-1]) and twin.dtype == torch.bfloat16
    close(twin.float().transpose(1, 2), want_bf, 1.2e-2)


@pytest.mark.parametrize("B,N,P,S", [(2, 3000, 64, 64), (1, 3000, 37, 64), (3, 4000, 333, 64), (2, 3000, 500, 16)])
def test_fused_sa_stage_super_tiles(B, N, P, S, monkeypatch):
    """sad_mlp_opts.super_tiles = 2 forces two tiles per context and phase in the general kernel (what it runs for SA1
    at full size) on small shapes: CTAs with one tile, odd tile counts, a partial last tile."""
    from sad_b200 import mlp as M
    monkeypatch.setattr(M, "SUPER_TILES", [2])
    monkeypatch.setattr(M, "FAST_SA", [False])
    rng = np.random.default_rng(N + P + S)
    xyz = (rng.random((B, N, 3), dtype=np.float32) * 3).astype(np.float32)
    feat = rng.standard_normal((B, 1, N)).astype(np.float32)
    inds = C.furthest_point_sample(xyz, P)
    new_xyz = np.stack([xyz[b][inds[b]] for b in range(B)])
    idx = C.ball_query(0.5, S, xyz, new_xyz)
    layers = make_layers(rng, [4, 64, 64, 128])
    x = O.query_and_group(xyz, new_xyz, feat, idx, np.float32(0.5), True, True)
    want = O.shared_mlp(x, layers, pool=True)
    got = M.sa_group_mlp(cu(xyz), cu(new_xyz), cu(feat), cu(idx), 0.5, M.prepare_layers(tlayers(layers)))
    close(got, want, 2e-2)



@pytest.mark.parametrize("B,n,C_,chans,last_relu", [
    (2, 1024, 256, [256, 256, 259], False),     # voting stack, linear last layer, c_last = 259
    (1, 200, 64, [64, 128], True),
    (3, 130, 100, [128, 64, 32], True),         # input width padded to 128
])
def test_fused_pointwise(B, n, C_, chans, last_relu):
    from sad_b200 import mlp as M
    rng = np.random.default_rng(n + C_)
    x = rng.standard_normal((B, C_, n)).astype(np.float32)
    layers = make_layers(rng, [C_] + chans)
    want = O.shared_mlp(x[..., None], layers, pool=False, last_relu=last_relu)[..., 0]
    want_bf = O.shared_mlp(x[..., None], layers, pool=False, last_relu=last_relu, emulate_bf16=True)[..., 0]
    mlp = M.prepare_layers(tlayers(layers))
    got = M.pointwise_mlp(cu(x), mlp, last_relu=last_relu)
    assert tuple(got.shape) == (B, chans[-1], n)
    close(got, want, 2e-2)
    close(got, want_bf, 8e-3)


@pytest.mark.parametrize("B,n,m,C2,C1,chans", [(2, 512, 256, 256, 256, [256, 256]), (1, 333, 50, 64, 128, [128, 64]),
                                               (2, 100, 20, 128, 0, [64, 64])])
def test_fused_fp_stage(B, n, m, C2, C1, chans):
    from sad_b200 import mlp as M
    rng = np.random.default_rng(n + m)
    unknown = (rng.random((B, n, 3), dtype=np.float32) * 2).astype(np.float32)
    known = (rng.random((B, m, 3), dtype=np.float32) * 2).astype(np.float32)
    kf = rng.standard_normal((B, C2, m)).astype(np.float32)
    uf = rng.standard_normal((B, C1, n)).astype(np.float32) if C1 else None
    layers = make_layers(rng, [C2 + C1] + chans)
    want = O.fp_module(unknown, known, uf, kf, layers)
    dist, idx = C.three_nn(unknown, known)
    w = O.interpolation_weights(dist)
    mlp = M.prepare_layers(tlayers(layers))
    got = M.fp_interp_mlp(cu(kf), cu(uf) if C1 else None, cu(idx), cu(w), mlp)
    assert tuple(got.shape) == (B, chans[-1], n)
    close(got, want, 2e-2)


def test_cf_to_cl_and_interp_cl_kernels():
    from sad_b200 import mlp as M
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 70, 333)).astype(np.float32)
    cl = M.to_cl_bf16(cu(x))
    assert tuple(cl.shape) == (2, 333, 70)
    np.testing.assert_array_equal(cl.float().cpu().numpy(), O.bf16_round(x.transpose(0, 2, 1)))
    padded = M.to_cl_bf16(cu(x), pad_to=128)
    assert tuple(padded.shape) == (2, 333, 128) and float(padded[:, :, 70:].abs().max()) == 0.0


def test_weight_image_matches_numpy_swizzle():
    """The packed image is byte-for-byte the K-major SWIZZLE_128B layout (host function, run here too)."""
    from sad_b200 import mlp as M
    rng = np.random.default_rng(1)
    W = rng.standard_normal((80, 150)).astype(np.float32)
    mlp = M.prepare_layers([(torch.from_numpy(W).to(DEV), torch.zeros(80, device=DEV)),
                            (torch.zeros(8, 80, device=DEV), torch.zeros(8, device=DEV))])
    lay = M.Layout(c0=192, c0_cols=range(150))
    imgs, *_ = mlp.packed(lay, 1)
    img = imgs[0].cpu().numpy().view(np.uint16)
    ref = (O.bf16_round(W).view(np.uint32) >> 16).astype(np.uint16)
    for (r, k) in [(0, 0), (5, 9), (79, 149), (33, 64), (8, 127), (17, 63)]:
        kc, kk = divmod(k, 64)
        unit = kk >> 3
        byte = kc * (80 * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((unit ^ (r & 7)) << 4) + (kk & 7) * 2
        assert img[byte // 2] == ref[r, k], (r, k)


def test_unfusable_shapes_are_rejected_not_routed_to_a_library_matmul():
    """hidden width not a multiple of 64 / four layers / nsample not a power of two -> SAD_EUNSUPPORTED on the host side."""
    from sad_b200 import mlp as M
    rng = np.random.default_rng(0)
    xyz = torch.rand(1, 256, 3, device=DEV)
    feat = torch.randn(1, 5, 256, device=DEV)
    idx = torch.zeros(1, 16, 16, dtype=torch.int32, device=DEV)
    idx24 = torch.zeros(1, 16, 24, dtype=torch.int32, device=DEV)
    odd = M.prepare_layers(tlayers(make_layers(rng, [8, 48, 64])))
    deep = M.prepare_layers(tlayers(make_layers(rng, [8, 64, 64, 64, 64])))
    good = M.prepare_layers(tlayers(make_layers(rng, [8, 64, 64])))
    for mlp, ix in ((odd, idx), (deep, idx), (good, idx24)):
        with pytest.raises(M.UnsupportedShape, match="SAD_EUNSUPPORTED"):
            M.sa_group_mlp(xyz, xyz[:, :16].contiguous(), feat, ix, 0.5, mlp)
    with pytest.raises(M.UnsupportedShape):
        M.pointwise_mlp(torch.randn(1, 8, 128, device=DEV), odd)
    with pytest.raises(M.UnsupportedShape):
        M.fp_interp_mlp(torch.randn(1, 8, 16, device=DEV), None, idx[:, :, :3].contiguous(), torch.rand(1, 16, 3, device=DEV), deep)
