"""Op parity: CUDA kernels (through the C ABI / autograd surface) vs the CPU oracle.

Indices bit-exact; three_interpolate / three_nn distances bit-exact as well (the
kernels follow the no-FMA arithmetic contract); scatter-add backward within 1e-5 of
the magnitude (atomics reorder the sum).  SURVEY.md section 4 tiers 2-3."""
import numpy as np
import pytest
import torch

from oracle import sad_oracle as O
from oracle import c_port as C

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def scene(rng, B, N, quant=None, kind="uniform"):
    if kind == "uniform":
        p = rng.random((B, N, 3), dtype=np.float32) * np.array([8, 8, 3], np.float32) - np.array([4, 4, 0], np.float32)
    else:  # clustered blobs: dense balls, lots of early exit
        c = rng.random((B, 8, 3), dtype=np.float32) * 4
        p = np.stack([c[b][rng.integers(0, 8, N)] for b in range(B)])
        p = p + rng.standard_normal((B, N, 3)).astype(np.float32) * np.float32(0.15)
    if quant:
        p = np.round(p / np.float32(quant)) * np.float32(quant)
    return np.ascontiguousarray(p, dtype=np.float32)


@pytest.fixture(scope="module")
def ops():
    import sad_b200
    return sad_b200


# ------------------------------------------------------------------------ FPS
@pytest.mark.parametrize("B,N,npoint,quant", [
    (1, 1, 1, None), (2, 5, 9, None), (3, 33, 7, None), (2, 255, 64, 0.5), (2, 256, 256, None),
    (3, 777, 100, None), (2, 1024, 512, 0.25), (2, 1025, 128, None), (4, 2048, 1024, None),
    (2, 4096, 300, None), (2, 4097, 200, None), (3, 9000, 257, 0.25), (2, 20000, 512, None),
])
def test_fps_matches_oracle(ops, B, N, npoint, quant):
    rng = np.random.default_rng(N * 31 + npoint)
    xyz = scene(rng, B, N, quant)
    want = C.furthest_point_sample(xyz, npoint)
    got = ops.furthest_point_sample(cu(xyz), npoint)
    assert got.dtype == torch.int32 and tuple(got.shape) == (B, npoint)
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("cs", [1, 2, 4, 8, 16])
def test_fps_every_cluster_size_is_bitexact(ops, cs):
    from sad_b200 import _lib
    rng = np.random.default_rng(cs)
    xyz = scene(rng, 3, 6000, 0.125)          # lattice => many exact ties
    want = C.furthest_point_sample(xyz, 300)
    got = ops.furthest_point_sample(cu(xyz), 300, None, "latency", False, cs)
    assert np.array_equal(got.cpu().numpy(), want)


def test_fps_duplicates_and_full_size(ops):
    p = np.ones((2, 300, 3), np.float32)
    assert ops.furthest_point_sample(cu(p), 5).cpu().tolist() == [[0] * 5] * 2
    rng = np.random.default_rng(5)
    xyz = scene(rng, 8, 40000)                                  # BASELINE config 2 shape
    want = C.furthest_point_sample(xyz, 2048)
    got = ops.furthest_point_sample(cu(xyz), 2048).cpu().numpy()
    assert np.array_equal(got, want)


def test_fps_large_scene_200k(ops):
    rng = np.random.default_rng(6)
    xyz = scene(rng, 1, 200000)
    want = C.furthest_point_sample(xyz, 256)
    assert np.array_equal(ops.furthest_point_sample(cu(xyz), 256).cpu().numpy(), want)
    with pytest.raises(RuntimeError):
        ops.furthest_point_sample(torch.zeros(1, 300000, 3, device=DEV), 4)


# ----------------------------------------------------------------- ball query
@pytest.mark.parametrize("B,N,npoint,radius,nsample,kind,quant", [
    (1, 1, 1, 0.5, 1, "uniform", None), (2, 50, 7, 0.8, 4, "uniform", None),
    (2, 777, 65, 0.6, 16, "uniform", 0.25), (3, 2048, 512, 0.4, 32, "blobs", None),
    (2, 2049, 130, 0.4, 64, "blobs", None), (2, 5000, 333, 0.3, 33, "uniform", None),
    (2, 4096, 64, 100.0, 64, "uniform", None), (2, 4100, 64, 1e-6, 8, "uniform", None),
    (1, 20001, 256, 0.5, 64, "blobs", None),     # scene base / tail not 16-byte aligned
    (3, 1333, 99, 0.7, 5, "uniform", 1.0),
])
def test_ball_query_matches_oracle(ops, B, N, npoint, radius, nsample, kind, quant):
    rng = np.random.default_rng(N + npoint)
    xyz = scene(rng, B, N, quant, kind)
    q = np.stack([xyz[b][rng.integers(0, N, npoint)] for b in range(B)])
    q[:, ::3] += np.float32(0.05)
    want = C.ball_query(radius, nsample, xyz, q)
    got = ops.ball_query(radius, nsample, cu(xyz), cu(q))
    assert got.dtype == torch.int32
    assert np.array_equal(got.cpu().numpy(), want)
    rt = (rng.random((B, npoint), dtype=np.float32) * np.float32(radius) + np.float32(0.05)).astype(np.float32)
    want = C.ball_query_adaptive(rt, nsample, xyz, q)
    got = ops.ball_query_adaptive(cu(rt), nsample, cu(xyz), cu(q))
    assert np.array_equal(got.cpu().numpy(), want)


def test_ball_query_known_answers_and_numpy_oracle(ops):
    xyz = np.zeros((1, 5, 3), np.float32)
    xyz[0, :, 0] = [0, .5, 1, 1.5, 2]
    q = np.array([[[0, 0, 0], [10, 0, 0], [1, 0, 0]]], np.float32)
    got = ops.ball_query(1.0, 4, cu(xyz), cu(q)).cpu().tolist()
    assert got == [[[0, 1, 0, 0], [0, 0, 0, 0], [1, 2, 3, 1]]]
    rng = np.random.default_rng(1)
    xyz = scene(rng, 2, 600, 0.5)
    q = xyz[:, :40].copy()
    assert np.array_equal(ops.ball_query(0.5, 8, cu(xyz), cu(q)).cpu().numpy(), O.ball_query(0.5, 8, xyz, q))


def test_ball_query_full_size_sa1(ops):
    rng = np.random.default_rng(9)
    xyz = scene(rng, 2, 40000)
    inds = C.furthest_point_sample(xyz, 2048)
    q = np.stack([xyz[b][inds[b]] for b in range(2)])
    want = C.ball_query(0.2, 64, xyz, q)
    assert np.array_equal(ops.ball_query(0.2, 64, cu(xyz), cu(q)).cpu().numpy(), want)


# ------------------------------------------------------------------- three_nn
@pytest.mark.parametrize("B,n,m,quant", [(1, 1, 3, None), (2, 100, 3, None), (2, 333, 40, 0.5),
                                         (3, 1024, 512, None), (2, 512, 256, 1.0), (1, 700, 5000, None)])
def test_three_nn_matches_oracle(ops, B, n, m, quant):
    rng = np.random.default_rng(n + m)
    u, k = scene(rng, B, n, quant), scene(rng, B, m, quant)
    wd, wi = C.three_nn(u, k)
    gd, gi = ops.three_nn(cu(u), cu(k))
    assert np.array_equal(gi.cpu().numpy(), wi)
    assert np.array_equal(gd.cpu().numpy(), wd)          # sqrt is correctly rounded on both sides
    with pytest.raises(ValueError):
        ops.three_nn(cu(u), cu(k[:, :2]))


# ------------------------------------------------- gather / group / interpolate
@pytest.mark.parametrize("B,C_,N,P,S", [(1, 1, 1, 1, 1), (2, 3, 100, 7, 5), (2, 4, 2048, 128, 64),
                                        (3, 17, 500, 33, 3), (2, 131, 2048, 256, 32), (1, 259, 512, 64, 16)])
def test_grouping_and_gather_fwd_bwd(ops, B, C_, N, P, S):
    rng = np.random.default_rng(B * 1000 + C_)
    f = rng.standard_normal((B, C_, N)).astype(np.float32)
    idx = rng.integers(0, N, (B, P, S)).astype(np.int32)
    ft = cu(f).requires_grad_(True)
    out = ops.grouping_operation(ft, cu(idx))
    assert np.array_equal(out.detach().cpu().numpy(), O.grouping_operation(f, idx))
    go = rng.standard_normal((B, C_, P, S)).astype(np.float32)
    out.backward(cu(go))
    want = O.grouping_operation_grad(go, idx, N)
    np.testing.assert_allclose(ft.grad.cpu().numpy(), want, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(want).max()))
    # gather (nsample == 1 surface)
    gi = idx[:, :, 0].copy()
    ft2 = cu(f).requires_grad_(True)
    out2 = ops.gather_operation(ft2, cu(gi))
    assert np.array_equal(out2.detach().cpu().numpy(), O.gather_operation(f, gi))
    go2 = rng.standard_normal((B, C_, P)).astype(np.float32)
    out2.backward(cu(go2))
    want2 = O.gather_operation_grad(go2, gi, N)
    np.testing.assert_allclose(ft2.grad.cpu().numpy(), want2, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(want2).max()))


@pytest.mark.parametrize("B,C_,m,n", [(1, 1, 3, 1), (2, 5, 40, 333), (2, 256, 256, 512), (3, 64, 512, 1024),
                                      (2, 19, 100, 1026),
                                      # point-major staged kernel: channel tails, n % 16 != 0, both chunk widths, largest m
                                      (2, 6, 384, 1000), (1, 70, 500, 1000), (2, 40, 1408, 2816), (1, 33, 1536, 3072)])
def test_three_interpolate_fwd_bwd(ops, B, C_, m, n):
    rng = np.random.default_rng(m + n)
    f = rng.standard_normal((B, C_, m)).astype(np.float32)
    idx = rng.integers(0, m, (B, n, 3)).astype(np.int32)
    w = O.interpolation_weights(rng.random((B, n, 3), dtype=np.float32))
    ft = cu(f).requires_grad_(True)
    out = ops.three_interpolate(ft, cu(idx), cu(w))
    assert np.array_equal(out.detach().cpu().numpy(), O.three_interpolate(f, idx, w))   # bit-exact (no FMA)
    go = rng.standard_normal((B, C_, n)).astype(np.float32)
    out.backward(cu(go))
    want = O.three_interpolate_grad(go, idx, w, m)
    np.testing.assert_allclose(ft.grad.cpu().numpy(), want, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(want).max()))


def test_gradcheck_small(ops):
    # finite differences in fp32 on tiny shapes (no fp64 kernels by design)
    rng = np.random.default_rng(0)
    f = cu(rng.standard_normal((1, 2, 6)).astype(np.float32)).requires_grad_(True)
    idx = cu(rng.integers(0, 6, (1, 3, 2)).astype(np.int32))
    out = ops.grouping_operation(f, idx)
    (out * out).sum().backward()
    f0 = f.detach().clone()
    num = torch.zeros_like(f0)
    eps = 1e-2
    for i in range(f0.numel()):
        d = torch.zeros_like(f0).view(-1)
        d[i] = eps
        fp = ops.grouping_operation((f0 + d.view_as(f0)).contiguous(), idx)
        fm = ops.grouping_operation((f0 - d.view_as(f0)).contiguous(), idx)
        num.view(-1)[i] = ((fp * fp).sum() - (fm * fm).sum()) / (2 * eps)
    torch.testing.assert_close(f.grad, num, rtol=1e-2, atol=1e-2)


# ------------------------------------------------------------ input validation
def test_bad_inputs_are_rejected_cleanly(ops):
    x = torch.zeros(2, 16, 3, device=DEV)
    with pytest.raises(RuntimeError):
        ops.furthest_point_sample(torch.zeros(2, 16, 3), 4)                  # CPU tensor
    with pytest.raises(ValueError):
        ops.furthest_point_sample(torch.zeros(2, 3, 16, device=DEV).transpose(1, 2), 4)   # non-contiguous
    with pytest.raises(TypeError):
        ops.furthest_point_sample(x.double(), 4)
    with pytest.raises(ValueError):
        ops.ball_query(0.5, 0, x, x)
    with pytest.raises(TypeError):
        ops.grouping_operation(torch.zeros(2, 4, 16, device=DEV), torch.zeros(2, 4, 4, device=DEV, dtype=torch.int64))
    # index outputs carry no gradient
    xr = x.clone().requires_grad_(True)
    assert not ops.furthest_point_sample(xr, 4).requires_grad
    assert not ops.ball_query(0.5, 4, xr, xr).requires_grad


def test_runs_on_the_current_stream(ops):
    rng = np.random.default_rng(2)
    xyz = scene(rng, 2, 3000)
    want = C.furthest_point_sample(xyz, 64)
    s = torch.cuda.Stream()
    x = cu(xyz)
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        got = ops.furthest_point_sample(x, 64)
    s.synchronize()
    assert np.array_equal(got.cpu().numpy(), want)


def _fps_order(xyz, npoint):
    inds = C.furthest_point_sample(xyz, npoint)
    return np.stack([xyz[b][inds[b]] for b in range(xyz.shape[0])])


@pytest.mark.parametrize("case", ["random", "duplicates", "lattice", "mixed"])
def test_fps_prefix_ordered_matches_sampler(case):
    """SURVEY H3 side note / VERDICT r1 item 5e: sampling FPS-ordered input is the identity unless a pick duplicates an
    earlier one; the device-side guard must send exactly those scenes to the real sampler.  Bit-exact vs the oracle."""
    from sad_b200 import ops
    rng = np.random.default_rng(7)
    if case == "random":
        X = rng.random((3, 6000, 3), dtype=np.float32) * 4
    elif case == "duplicates":            # 300 distinct points: picks past the 300th have min-distance 0
        base = rng.random((2, 300, 3), dtype=np.float32)
        X = base[:, rng.integers(0, 300, 2500)]
    elif case == "lattice":               # exact ties everywhere, no duplicates
        g = np.stack(np.meshgrid(*[np.arange(18, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3)
        X = np.stack([g[rng.permutation(len(g))] * np.float32(0.25) for _ in range(2)])
    else:                                 # one scene passes the guard, one does not
        a = rng.random((1, 2500, 3), dtype=np.float32)
        base = rng.random((1, 200, 3), dtype=np.float32)
        X = np.concatenate([a, base[:, rng.integers(0, 200, 2500)]], 0)
    Y = _fps_order(np.ascontiguousarray(X), 2048)
    for K in (1024, 512, 256, 1):
        want = C.furthest_point_sample(Y, K)
        got = ops.furthest_point_sample(cu(Y), K, None, "latency", True)
        np.testing.assert_array_equal(got.cpu().numpy(), want)
        Y = np.stack([Y[b][want[b]] for b in range(Y.shape[0])]) if K > 1 else Y
    if case == "random":
        assert (want >= 0).all()


def test_glue_kernels_match_their_torch_expressions():
    """the one-launch replacements of the torch glue are bit-equal to what they replace and to the
    oracle: new_xyz gather, three_nn + normalised weights, size -> radius."""
    from sad_b200 import ops
    rng = np.random.default_rng(11)
    xyz = (rng.random((3, 5000, 3), dtype=np.float32) * 5).astype(np.float32)
    inds = C.furthest_point_sample(xyz, 257)
    want_xyz = np.stack([xyz[b][inds[b]] for b in range(3)])
    got, got4 = ops.gather_points(cu(xyz), cu(inds), with_xyzw=True)
    assert np.array_equal(got.cpu().numpy(), want_xyz)
    assert np.array_equal(got4.cpu().numpy()[..., :3], want_xyz) and not got4.cpu().numpy()[..., 3].any()
    unknown = (rng.random((2, 700, 3), dtype=np.float32) * 4).astype(np.float32)
    known = (rng.random((2, 130, 3), dtype=np.float32) * 4).astype(np.float32)
    known[0, 5] = unknown[0, 9]                        # a zero distance: weight = 1 / 1e-8 normalised
    d, i, w = ops.three_nn_weights(cu(unknown), cu(known))
    wd, wi = O.three_nn(unknown, known)
    assert np.array_equal(i.cpu().numpy(), wi) and np.array_equal(d.cpu().numpy(), wd)
    assert np.array_equal(w.cpu().numpy(), O.interpolation_weights(wd))
    size = (rng.random((4, 256, 3), dtype=np.float32) * 3).astype(np.float32)
    for (alpha, lo, hi) in ((1.0, 0.1, 1.2), (0.7, 0.05, 0.9)):
        r = ops.size_to_radius(cu(size), alpha, lo, hi)
        assert np.array_equal(r.cpu().numpy(), O.size_to_radius(size, alpha, lo, hi))


# ----------------------------------------------------------------------------- deterministic scatter-add
@pytest.fixture
def deterministic(ops):
    prev = ops.set_deterministic(True)
    yield ops
    ops.set_deterministic(prev)


@pytest.mark.parametrize("B,C_,N,P,S_", [(2, 19, 500, 64, 16), (1, 4, 20000, 2048, 64), (3, 64, 1024, 333, 8),
                                          (1, 3, 60000, 128, 32),       # N beyond the shared-memory cursor array
                                          (1, 5, 16, 200, 64)])         # every destination hit hundreds of times
def test_deterministic_grouping_backward_is_bit_exact(deterministic, B, C_, N, P, S_):
    """The sort-by-destination backward adds each destination's addends in ascending source position -- the order of the
    oracle's np.add.at -- so it equals the oracle bit for bit and itself from run to run; the atomic kernel only agrees
    to rounding."""
    ops = deterministic
    rng = np.random.default_rng(N + P)
    idx = rng.integers(0, N, (B, P, S_)).astype(np.int32)
    idx[:, :, S_ // 2:] = idx[:, :, :1]                    # first-hit padding: heavy duplication inside a ball
    go = rng.standard_normal((B, C_, P, S_)).astype(np.float32)
    want = O.grouping_operation_grad(go, idx, N)
    f = torch.zeros(B, C_, N, device=DEV, requires_grad=True)
    runs = []
    for _ in range(3):
        f.grad = None
        ops.grouping_operation(f, cu(idx)).backward(cu(go))
        runs.append(f.grad.cpu().numpy().copy())
    assert np.array_equal(runs[0], want)
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    # gather_operation shares the kernels
    gi = rng.integers(0, N, (B, P)).astype(np.int32)
    go2 = rng.standard_normal((B, C_, P)).astype(np.float32)
    f.grad = None
    ops.gather_operation(f, cu(gi)).backward(cu(go2))
    assert np.array_equal(f.grad.cpu().numpy(), O.gather_operation_grad(go2, gi, N))


@pytest.mark.parametrize("B,C_,m,n", [(2, 5, 40, 333), (2, 256, 256, 512), (1, 70, 3, 1000), (1, 8, 512, 1024)])
def test_deterministic_three_interpolate_backward_is_bit_exact(deterministic, B, C_, m, n):
    ops = deterministic
    rng = np.random.default_rng(m + n)
    f = cu(rng.standard_normal((B, C_, m)).astype(np.float32)).requires_grad_(True)
    idx = rng.integers(0, m, (B, n, 3)).astype(np.int32)
    w = O.interpolation_weights(rng.random((B, n, 3), dtype=np.float32))
    go = rng.standard_normal((B, C_, n)).astype(np.float32)
    want = O.three_interpolate_grad(go, idx, w, m)
    for _ in range(2):
        f.grad = None
        ops.three_interpolate(f, cu(idx), cu(w)).backward(cu(go))
        assert np.array_equal(f.grad.cpu().numpy(), want)


def test_scatter_plan_skips_out_of_range_indices(ops):
    """indices outside [0, N) never reach memory in the deterministic path; they are dropped from the plan."""
    import ctypes
    from sad_b200 import _lib
    lib = _lib.load()
    vp = ctypes.c_void_p
    B, N, PS = 1, 10, 64
    idx = torch.arange(PS, dtype=torch.int32, device=DEV).view(B, PS) - 20          # -20 .. 43: 10 valid
    order = torch.full((B, PS), -7, dtype=torch.int32, device=DEV)
    offsets = torch.empty((B, N + 1), dtype=torch.int32, device=DEV)
    st = vp(torch.cuda.current_stream().cuda_stream)
    assert lib.sad_scatter_plan_build(B, N, PS, vp(idx.data_ptr()), vp(order.data_ptr()), vp(offsets.data_ptr()), st) == 0
    assert offsets.cpu().tolist() == [list(range(11))]
    assert order[0, :10].cpu().tolist() == list(range(20, 30)) and int((order[0, 10:] != -7).sum()) == 0


def test_out_of_range_indices_never_touch_memory_out_of_bounds(ops):
    """user-supplied idx is untrusted.  Forward gathers clamp into [0, N) (negative -> N - 1), backward
    scatters skip; the in-range part of the result is unaffected."""
    rng = np.random.default_rng(0)
    B, C_, N, P, S_ = 2, 20, 300, 64, 16
    f = rng.standard_normal((B, C_, N)).astype(np.float32)
    idx = rng.integers(0, N, (B, P, S_)).astype(np.int32)
    bad = idx.copy()
    bad[0, 0, 0], bad[1, 3, 5], bad[1, 7, 1] = N, -5, 2 ** 31 - 1
    ref = idx.copy()
    ref[0, 0, 0], ref[1, 3, 5], ref[1, 7, 1] = N - 1, N - 1, N - 1
    ft = cu(f).requires_grad_(True)
    out = ops.grouping_operation(ft, cu(bad))
    assert np.array_equal(out.detach().cpu().numpy(), O.grouping_operation(f, ref))
    go = rng.standard_normal((B, C_, P, S_)).astype(np.float32)
    out.backward(cu(go))
    go_ref = go.copy()
    go_ref[0, :, 0, 0] = 0
    go_ref[1, :, 3, 5] = 0
    go_ref[1, :, 7, 1] = 0
    want = O.grouping_operation_grad(go_ref, idx, N)
    np.testing.assert_allclose(ft.grad.cpu().numpy(), want, rtol=1e-5, atol=1e-5 * np.abs(want).max())
    # three_interpolate: same policy
    m, n = 40, 256
    kf = rng.standard_normal((B, C_, m)).astype(np.float32)
    ii = rng.integers(0, m, (B, n, 3)).astype(np.int32)
    w = O.interpolation_weights(rng.random((B, n, 3), dtype=np.float32))
    ib = ii.copy()
    ib[0, 0, 0], ib[1, 9, 2] = m + 7, -1
    ir = ii.copy()
    ir[0, 0, 0], ir[1, 9, 2] = m - 1, m - 1
    kt = cu(kf).requires_grad_(True)
    o2 = ops.three_interpolate(kt, cu(ib), cu(w))
    assert np.array_equal(o2.detach().cpu().numpy(), O.three_interpolate(kf, ir, w))
    o2.backward(torch.ones_like(o2))
    assert torch.isfinite(kt.grad).all()


def test_scene_grid_goes_stale_when_xyz_is_written_in_place(ops):
    from sad_b200 import ops as _ops
    xyz = torch.rand(1, 9000, 3, device=DEV) * 4
    grid = _ops.build_scene_grid(xyz)
    _ops.furthest_point_sample(xyz, 64, grid)
    xyz.mul_(0.5)
    with pytest.raises(ValueError, match="stale"):
        _ops.furthest_point_sample(xyz, 64, grid)
