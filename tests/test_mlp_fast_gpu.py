"""Shape-specialised fused SA kernel (csrc/mlp_sa.cu) vs the oracle and vs the general kernel.

Every instance the library compiles is exercised: SA1 (one CTA, 4 tile contexts, special K step only), SA2 on one
CTA and on a CTA pair (TMA gather4, cta_group::2), SA3/SA4 and the vote aggregation (zero-padded outputs,
per-cluster radius) on CTA pairs; small cases (partial tiles, an odd number of tiles for the pair kernels) and cases
deep enough that every CTA recycles its tile buffers and contexts several times."""
import numpy as np
import pytest
import torch

from oracle import sad_oracle as O
from oracle import c_port as C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def close(got, want, tol, floor_frac=0.05, rel=None):
    got = got.detach().float().cpu().numpy()
    scale = max(1e-6, float(np.abs(want).max()))
    err = float(np.abs(got - want).max())
    assert err <= tol * scale, f"max abs err {err:.4g} vs scale {scale:.4g} (tol {tol})"
    if rel is not None:      # element-wise relative error above a magnitude floor
        m = np.abs(want) > floor_frac * scale
        r = float((np.abs(got - want)[m] / np.abs(want)[m]).max()) if m.any() else 0.0
        assert r <= rel, f"max element-wise relative err {r:.4g} above {floor_frac} of scale (tol {rel})"


def make_layers(rng, chans, bias_std=0.1):
    return [((rng.standard_normal((co, ci)) / np.sqrt(ci)).astype(np.float32),
             (rng.standard_normal(co) * bias_std).astype(np.float32)) for ci, co in zip(chans[:-1], chans[1:])]


CASES = [
    # name, mode, B, N, P, S, Cf, hidden, radius, adaptive
    ("sa1-small", True, 1, 3000, 64, 64, 1, [64, 64, 128], 0.5, False),
    ("sa1-odd-tiles", True, 1, 3000, 1, 64, 1, [64, 64, 128], 0.5, False),          # half a tile
    ("sa1-deep", True, 2, 20000, 2048, 64, 1, [64, 64, 128], 0.3, False),           # ~28 tiles per CTA
    ("sa2-single-small", "single", 2, 2048, 128, 32, 128, [128, 128, 256], 0.6, False),
    ("sa2-single-deep", "single", 4, 2048, 1024, 32, 128, [128, 128, 256], 0.6, False),
    ("sa2-pair-small", "pair", 2, 2048, 128, 32, 128, [128, 128, 256], 0.6, False),
    ("sa2-pair-odd", "pair", 1, 2048, 4, 32, 128, [128, 128, 256], 0.6, False),       # one tile: the pair's second CTA idles
    ("sa2-pair-deep", "pair", 4, 2048, 1024, 32, 128, [128, 128, 256], 0.6, False),
    ("sa3-pair-small", True, 1, 1024, 64, 16, 256, [128, 128, 256], 0.9, False),
    ("sa3-pair-deep", True, 8, 1024, 512, 16, 256, [128, 128, 256], 0.9, False),
    ("agg-pair", True, 2, 1024, 128, 16, 256, [128, 128, 128], 0.3, True),          # 128 outputs zero-padded, per-cluster radius
    ("agg-pair-deep", True, 8, 1024, 256, 16, 256, [128, 128, 128], 0.3, True),
]


@pytest.mark.parametrize("name,mode,B,N,P,S,Cf,hidden,radius,adaptive", CASES, ids=[c[0] for c in CASES])
def test_fast_sa_stage(name, mode, B, N, P, S, Cf, hidden, radius, adaptive):
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
    mlp = M.prepare_layers([(cu(W), cu(b)) for W, b in layers])
    args = (cu(xyz), cu(new_xyz), cu(feat), cu(idx), rad_g, mlp)
    saved = M.FAST_SA[0]
    try:
        M.FAST_SA[0] = mode
        layout = M.sa_layout(Cf, True)
        inst = M._fast_instance(mlp, layout, S, P)
        assert inst >= 0, "no specialised instance picked this shape"
        got = M.sa_group_mlp(*args, use_xyz=True, normalize_xyz=True)
        torch.cuda.synchronize()
        got2 = M.sa_group_mlp(*args, use_xyz=True, normalize_xyz=True)      # scheduler words were re-armed by launch 1
        torch.cuda.synchronize()
        M.FAST_SA[0] = False
        ref = M.sa_group_mlp(*args, use_xyz=True, normalize_xyz=True)       # general kernel
    finally:
        M.FAST_SA[0] = saved
    assert tuple(got.shape) == (B, hidden[-1], P) and got.dtype == torch.float32
    close(got, want, 2e-2)
    close(got, want_bf, 8e-3, rel=4e-2)
    assert torch.equal(got, got2), "second launch on the same scheduler words differs"
    close(got, ref.detach().float().cpu().numpy(), 8e-3)
    twin = got._sad_cl
    assert tuple(twin.shape) == (B, P, hidden[-1]) and twin.dtype == torch.bfloat16
    close(twin.float().transpose(1, 2), want_bf, 1.2e-2)


@pytest.mark.parametrize("tpc", [1, 6])
@pytest.mark.parametrize("B,n,m", [(1, 128, 64), (2, 512, 256), (8, 1024, 512), (3, 384, 130)])
def test_fast_fp_stage(B, n, m, tpc, monkeypatch):
    """csrc/mlp_pw.cu kind 0: three_interpolate + concat + 2-layer MLP in one launch vs the oracle's FP module and vs
    the general kernel (three_interpolate_cl + fused_mlp_kernel)."""
    from sad_b200 import mlp as M
    monkeypatch.setattr(M, "TILES_PER_CTA", [tpc])          # 6 -> two tiles per CTA: the ring wraps across tiles
    rng = np.random.default_rng(n + m)
    unknown = (rng.random((B, n, 3), dtype=np.float32) * 4).astype(np.float32)
    known = (rng.random((B, m, 3), dtype=np.float32) * 4).astype(np.float32)
    kf = rng.standard_normal((B, 256, m)).astype(np.float32)
    uf = rng.standard_normal((B, 256, n)).astype(np.float32)
    layers = make_layers(rng, [512, 256, 256])
    dist, idx = O.three_nn(unknown, known)
    w = O.interpolation_weights(dist)
    x = np.concatenate([O.three_interpolate(kf, idx, w), uf], axis=1)[:, :, :, None]
    want = O.shared_mlp(x, layers, pool=False)[..., 0]
    mlp = M.prepare_layers([(cu(W), cu(b)) for W, b in layers])
    args = (cu(kf), cu(uf), cu(idx), cu(w), mlp)
    saved = M.FAST_PW[0]
    try:
        M.FAST_PW[0] = True
        got = M.fp_interp_mlp(*args)
        M.FAST_PW[0] = False
        ref = M.fp_interp_mlp(*args)
    finally:
        M.FAST_PW[0] = saved
    assert tuple(got.shape) == (B, 256, n)
    close(got, want, 2e-2)
    close(got, ref.detach().float().cpu().numpy(), 8e-3)
    close(got._sad_cl.float().transpose(1, 2), want, 2e-2)


@pytest.mark.parametrize("tpc", [1, 6])
@pytest.mark.parametrize("B,n", [(1, 128), (8, 1024), (2, 384)])
def test_fast_voting_stage(B, n, tpc, monkeypatch):
    """csrc/mlp_pw.cu kind 1: 3-layer voting MLP fused with vote = seed + y vs the oracle's voting module."""
    from sad_b200 import mlp as M
    monkeypatch.setattr(M, "TILES_PER_CTA", [tpc])
    rng = np.random.default_rng(n)
    seed_xyz = (rng.random((B, n, 3), dtype=np.float32) * 4).astype(np.float32)
    sf = rng.standard_normal((B, 256, n)).astype(np.float32)
    layers = make_layers(rng, [256, 256, 256, 259])
    want_xyz, want_feat = O.voting_module(seed_xyz, sf, layers)
    mlp = M.prepare_layers([(cu(W), cu(b)) for W, b in layers])
    assert M.vote_fast_ok(cu(sf), mlp)
    vx, vf = M.vote_mlp_fast(cu(seed_xyz), cu(sf), mlp)
    close(vx, want_xyz, 2e-2)
    close(vf, want_feat, 2e-2)
    close(vf._sad_cl.float().transpose(1, 2), want_feat, 2e-2)
    # the offsets alone (vote - seed): the bf16 bar at the scale of the MLP output, not of the coordinates
    y = O.shared_mlp(sf[..., None], layers, pool=False, last_relu=False)[..., 0]
    close(vx - cu(seed_xyz), y[:, :3, :].transpose(0, 2, 1), 2e-2)


# ----------------------------------------------------------------------------- duplicate-free SA stages
def _ball_idx(rng, B, N, P, S, radius):
    xyz = (rng.random((B, N, 3), dtype=np.float32) * 3).astype(np.float32)
    inds = C.furthest_point_sample(xyz, P)
    new_xyz = np.stack([xyz[b][inds[b]] for b in range(B)])
    return xyz, new_xyz, C.ball_query(radius, S, xyz, new_xyz)


@pytest.mark.parametrize("which,B,N,P,radius,tpc,slot", [
    ("sa1", 2, 6000, 256, 0.2, 1, 16),        # sparse balls: mostly 1- and 2-slot runs
    ("sa1", 1, 6000, 128, 0.6, 1, 16),        # dense balls: every point needs all 64 samples (4-slot runs)
    ("sa1", 3, 9000, 512, 0.3, 6, 0),         # mixed, narrow grid
    ("sa2", 2, 2048, 256, 0.25, 1, 16),       # nsample 32, 16-sample slots: 1- and 2-slot runs
    ("sa2", 2, 2048, 128, 1.5, 1, 16),        # all 2-slot runs
    ("sa2", 1, 2048, 64, 0.02, 6, 16),        # nearly empty balls: 1 slot per point, a single partial tile
    ("sa2", 1, 2048, 64, 0.02, 6, 0),         # 1 slot per point (automatic slot size), a single partial tile
    ("sa2", 4, 2048, 1024, 0.4, 1, 0),        # the benchmark's shape and radius
])
def test_duplicate_free_stage_is_bit_identical(which, B, N, P, radius, tpc, slot, monkeypatch):
    """sad_sa_mlp_dedup_fwd (plan kernel + 16-sample sibling instance over slots) against the ordinary launch of the same
    stage: identical bits in both output layouts, and both within the bf16 bar of the oracle."""
    from sad_b200 import mlp as M
    monkeypatch.setattr(M, "TILES_PER_CTA", [tpc])
    monkeypatch.setattr(M, "DEDUP_NSAMPLE", (32, 64))      # (nsample 64 is off by default: it does not pay there)
    monkeypatch.setattr(M, "DEDUP_SLOT", [slot])
    monkeypatch.setattr(M, "_DEDUP_OK", {})
    rng = np.random.default_rng(P + int(radius * 100))
    S, Cf, hidden = (64, 1, [64, 64, 128]) if which == "sa1" else (32, 128, [128, 128, 256])
    xyz, new_xyz, idx = _ball_idx(rng, B, N, P, S, radius)
    feat = rng.standard_normal((B, Cf, N)).astype(np.float32)
    layers = make_layers(rng, [Cf + 3] + hidden)
    mlp = M.prepare_layers([(cu(W), cu(b)) for W, b in layers])
    outs = {}
    for dedup in (False, True):
        monkeypatch.setattr(M, "DEDUP_SA", [dedup])
        got = M.sa_group_mlp(cu(xyz), cu(new_xyz), cu(feat), cu(idx), radius, mlp, use_xyz=True, normalize_xyz=True)
        outs[dedup] = (got.cpu().numpy().copy(), got._sad_cl.float().cpu().numpy().copy())
    assert np.array_equal(outs[True][0], outs[False][0]) and np.array_equal(outs[True][1], outs[False][1])
    want = O.shared_mlp(O.query_and_group(xyz, new_xyz, feat, idx, np.float32(radius), True, True), layers, pool=True)
    close(torch.from_numpy(outs[True][0]), want, 2e-2)
    # the slot statistics the plan is built on: how much of the stage is padding
    first = idx[..., :1]
    dup_tail16 = float((idx[..., 16:] == first).all(-1).mean())
    assert 0.0 <= dup_tail16 <= 1.0


def test_duplicate_free_stage_handles_arbitrary_indices(monkeypatch):
    """idx that is not a ball query's (no padding structure, duplicates in the middle, per-cluster radius): the plan
    checks the tail per point, so the result is still exact."""
    from sad_b200 import mlp as M
    monkeypatch.setattr(M, "_DEDUP_OK", {})
    rng = np.random.default_rng(3)
    B, N, P, S = 2, 3000, 128, 32
    xyz = (rng.random((B, N, 3), dtype=np.float32) * 3).astype(np.float32)
    new_xyz = xyz[:, :P].copy()
    idx = rng.integers(0, N, (B, P, S)).astype(np.int32)
    idx[0, :40, 16:] = idx[0, :40, :1]            # looks padded after 16 samples
    idx[0, 40:60, 20:] = idx[0, 40:60, :1]        # padded, but only from sample 20: needs both slots
    idx[1, :10, :] = idx[1, :10, :1]              # a single neighbour
    rad = (0.3 + rng.random((B, P))).astype(np.float32)
    feat = rng.standard_normal((B, 128, N)).astype(np.float32)
    layers = make_layers(rng, [131, 128, 128, 256])
    mlp = M.prepare_layers([(cu(W), cu(b)) for W, b in layers])
    outs = []
    for dedup in (False, True):
        monkeypatch.setattr(M, "DEDUP_SA", [dedup])
        outs.append(M.sa_group_mlp(cu(xyz), cu(new_xyz), cu(feat), cu(idx), cu(rad), mlp).cpu().numpy().copy())
    assert np.array_equal(outs[0], outs[1])
    want = O.shared_mlp(O.query_and_group(xyz, new_xyz, feat, idx, rad, True, True), layers, pool=True)
    close(torch.from_numpy(outs[1]), want, 2e-2)
