"""Scene grid (csrc/grid.cu), exact culled FPS (csrc/fps_cull.cu) and grid ball query vs the CPU
oracle: indices bit-exact, ties to the lowest original index, whatever the spatial sort did.
SURVEY.md section 8(f) rank 2 ("must remain bit-exact vs oracle")."""
import numpy as np
import pytest
import torch

from oracle import c_port as C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = 32
HDR, CELLS = 64, (G ** 3 + 1) * 4
CELL_BYTES = (CELLS + 15) // 16 * 16


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def scene(rng, B, N, quant=None, kind="uniform"):
    if kind == "uniform":
        p = rng.random((B, N, 3), dtype=np.float32) * np.array([8, 8, 3], np.float32) - np.array([4, 4, 0], np.float32)
    elif kind == "surface":
        from sad_b200.scenes import make_scenes
        p = make_scenes(B, N, "surface", first_scene=int(rng.integers(0, 1000)))[0]
    else:  # dense blobs
        c = rng.random((B, 8, 3), dtype=np.float32) * 4
        p = np.stack([c[b][rng.integers(0, 8, N)] for b in range(B)])
        p = p + rng.standard_normal((B, N, 3)).astype(np.float32) * np.float32(0.15)
    if quant:
        p = np.round(p / np.float32(quant)) * np.float32(quant)
    return np.ascontiguousarray(p, dtype=np.float32)


@pytest.fixture(scope="module")
def ops():
    import sad_b200
    from sad_b200 import ops as O
    return O


@pytest.mark.parametrize("B,N", [(1, 1), (2, 31), (3, 1000), (2, 40000)])
def test_grid_build_is_a_cell_sorted_permutation(ops, B, N):
    rng = np.random.default_rng(N)
    xyz = scene(rng, B, N)
    grid = ops.build_scene_grid(cu(xyz))
    raw = grid.workspace.cpu().numpy()
    stride = HDR + CELL_BYTES + N * 16 + (N * 4 + 15) // 16 * 16
    assert raw.size == B * stride
    for b in range(B):
        blk = raw[b * stride:(b + 1) * stride]
        hdr = blk[:16].view(np.float32)
        start = blk[HDR:HDR + CELLS].view(np.uint32)
        pts = blk[HDR + CELL_BYTES:HDR + CELL_BYTES + N * 16].view(np.float32).reshape(N, 4)
        oidx = pts[:, 3].copy().view(np.uint32)
        assert np.array_equal(np.sort(oidx), np.arange(N, dtype=np.uint32))          # a permutation ...
        assert np.array_equal(pts[:, :3], xyz[b][oidx])                              # ... carrying its coordinates
        assert np.array_equal(hdr[:3], xyz[b].min(axis=0))
        assert start[0] == 0 and start[-1] == N and np.all(np.diff(start.astype(np.int64)) >= 0)
        # every point lies in the cell range it was filed under (same fp32 cell function)
        t = np.floor((pts[:, :3] - hdr[:3]).astype(np.float32) * hdr[3]).astype(np.float32)
        c = np.clip(t, 0, G - 1).astype(np.int64)
        cell = c[:, 0] + G * (c[:, 1] + G * c[:, 2])
        assert np.all(np.diff(cell) >= 0)
        pos = np.arange(N)
        assert np.all(start[cell] <= pos) and np.all(pos < start[cell + 1])


@pytest.mark.parametrize("B,N,npoint,quant,kind", [
    (1, 1, 1, None, "uniform"), (2, 5, 9, None, "uniform"), (3, 33, 7, None, "uniform"),
    (2, 255, 64, 0.5, "uniform"), (2, 1000, 300, None, "blobs"), (4, 2048, 1024, None, "surface"),
    (2, 4097, 200, 0.25, "uniform"),                     # lattice: many exact ties
    (2, 10752, 128, None, "uniform"),                    # exactly one CTA's capacity
    (2, 10753, 128, None, "blobs"),                      # first size that needs a cluster of 2
    (2, 20000, 512, 0.25, "uniform"),                    # cluster of 2, ties
    (8, 40000, 2048, None, "surface"),                   # BASELINE config 2 shape (cluster of 4)
    (1, 50000, 300, None, "uniform"),                    # cluster of 8
    (1, 100003, 200, None, "blobs"),                     # N % 32 != 0, 7 bucket slots per lane
    (1, 250000, 64, None, "uniform"),                    # beyond the cluster capacity: single-CTA kernel
])
def test_culled_fps_matches_oracle(ops, B, N, npoint, quant, kind):
    rng = np.random.default_rng(N * 7 + npoint)
    xyz = scene(rng, B, N, quant, kind)
    want = C.furthest_point_sample(xyz, npoint)
    x = cu(xyz)
    got = ops.furthest_point_sample(x, npoint, ops.build_scene_grid(x))
    assert got.dtype == torch.int32 and tuple(got.shape) == (B, npoint)
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("cs", [-1, 1, 2, 4, 16])
def test_culled_fps_every_kernel_variant_is_bitexact(ops, cs):
    from sad_b200 import _lib
    rng = np.random.default_rng(cs + 7)
    xyz = scene(rng, 2, 9000, 0.125)          # lattice => many exact ties
    want = C.furthest_point_sample(xyz, 300)
    x = cu(xyz)
    got = ops.furthest_point_sample(x, 300, ops.build_scene_grid(x), "latency", False, cs)
    assert np.array_equal(got.cpu().numpy(), want)


def test_culled_fps_duplicates_and_npoint_beyond_distinct(ops):
    p = np.ones((2, 300, 3), np.float32)
    x = cu(p)
    assert ops.furthest_point_sample(x, 5, ops.build_scene_grid(x)).cpu().tolist() == [[0] * 5] * 2
    # 40 distinct positions, each repeated; more picks than distinct points
    rng = np.random.default_rng(3)
    base = rng.random((40, 3), dtype=np.float32)
    p = np.ascontiguousarray(base[rng.integers(0, 40, (2, 900))])
    want = C.furthest_point_sample(p, 64)
    x = cu(p)
    assert np.array_equal(ops.furthest_point_sample(x, 64, ops.build_scene_grid(x)).cpu().numpy(), want)


def test_culled_fps_is_the_default_for_large_scenes_and_checks_its_grid(ops):
    rng = np.random.default_rng(9)
    xyz = scene(rng, 2, 9000)
    x = cu(xyz)
    want = C.furthest_point_sample(xyz, 100)
    assert np.array_equal(ops.furthest_point_sample(x, 100).cpu().numpy(), want)        # builds its own grid
    other = cu(scene(rng, 2, 9000))
    with pytest.raises(ValueError):
        ops.furthest_point_sample(other, 10, ops.build_scene_grid(x))


@pytest.mark.parametrize("B,N,npoint,radius,nsample,kind,quant", [
    (1, 1, 1, 0.5, 1, "uniform", None), (2, 50, 7, 0.8, 4, "uniform", None),
    (2, 777, 65, 0.6, 16, "uniform", 0.25), (3, 2048, 512, 0.4, 32, "blobs", None),
    (2, 5000, 333, 0.3, 33, "uniform", None),
    (2, 4096, 64, 100.0, 64, "uniform", None),          # ball covers everything: overflow -> exact fallback
    (2, 4100, 64, 1e-6, 8, "uniform", None),            # ball covers nothing but the query itself
    (1, 20001, 256, 0.5, 64, "blobs", None),            # dense: thousands of hits per ball
    (3, 1333, 99, 0.7, 5, "uniform", 1.0),              # lattice
    (8, 40000, 2048, 0.2, 64, "surface", None),         # BASELINE config 2, SA1
    (2, 20000, 1024, 0.35, 16, "surface", None),
])
def test_grid_ball_query_matches_oracle(ops, B, N, npoint, radius, nsample, kind, quant):
    rng = np.random.default_rng(N + npoint)
    xyz = scene(rng, B, N, quant, kind)
    q = np.stack([xyz[b][rng.integers(0, N, npoint)] for b in range(B)])
    q[:, ::3] += np.float32(0.05)
    q[:, 1::7] += np.float32(30.0)                       # some queries far outside the scene
    x, qq = cu(xyz), cu(q)
    grid = ops.build_scene_grid(x)
    want = C.ball_query(radius, nsample, xyz, q)
    got = ops.ball_query(radius, nsample, x, qq, grid)
    assert got.dtype == torch.int32
    assert np.array_equal(got.cpu().numpy(), want)
    rt = (rng.random((B, npoint), dtype=np.float32) * np.float32(radius) + np.float32(0.05)).astype(np.float32)
    want = C.ball_query_adaptive(rt, nsample, xyz, q)
    got = ops.ball_query_adaptive(cu(rt), nsample, x, qq, grid)
    assert np.array_equal(got.cpu().numpy(), want)


def test_grid_ball_query_duplicates_overflow(ops):
    p = np.zeros((1, 5000, 3), np.float32)
    p[0, 2500:] = 1.0
    q = np.array([[[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5]]], np.float32)
    x = cu(p)
    got = ops.ball_query(0.1, 8, x, cu(q), ops.build_scene_grid(x)).cpu().numpy()
    assert np.array_equal(got, C.ball_query(0.1, 8, p, q))
    assert got[0, 0].tolist() == list(range(8)) and got[0, 1].tolist() == list(range(2500, 2508)) and not got[0, 2].any()


@pytest.mark.parametrize("N,npoint", [(9000, 300), (40000, 512), (70001, 128)])
def test_fps_scheduling_policy_never_changes_the_result(ops, N, npoint):
    """C ABI sad_furthest_point_sample_grid_policy_fwd: latency (cluster) and throughput (one SM per scene)."""
    rng = np.random.default_rng(N)
    xyz = (rng.random((2, N, 3), dtype=np.float32) * np.array([6, 6, 3], np.float32)).astype(np.float32)
    x = cu(xyz)
    grid = ops.build_scene_grid(x)
    want = C.furthest_point_sample(xyz, npoint)
    for policy in ("latency", "throughput"):
        got = ops.furthest_point_sample(x, npoint, grid, policy)
        assert np.array_equal(got.cpu().numpy(), want), policy
    with pytest.raises(KeyError):
        ops.furthest_point_sample(x, npoint, grid, "fastest")


@pytest.mark.parametrize("B,N,npoint,quant,kind", [
    (1, 1, 1, None, "uniform"), (2, 5, 9, None, "uniform"), (3, 33, 2, None, "uniform"), (2, 64, 3, None, "uniform"),
    (2, 255, 64, 0.5, "uniform"),                        # lattice: many exact ties, most picks duplicates at the end
    (2, 1000, 300, None, "blobs"), (4, 2048, 1024, None, "surface"),
    (2, 4097, 200, 0.25, "uniform"),                     # lattice
    (2, 20000, 512, 0.25, "uniform"),                    # two register sets of bucket state, ties
    (8, 40000, 2048, None, "surface"),                   # the benchmark's sampling (three register sets)
    (2, 40000, 2500, 0.05, "surface"),                   # more picks than the output ring holds, ties
    (1, 46000, 300, None, "blobs"),                      # the largest scene whose min-distances fit shared memory
])
def test_throughput_fps_matches_oracle(ops, B, N, npoint, quant, kind):
    """The one-SM-per-scene kernel (throughput policy) over its whole range of shapes and instances (<R register sets,
    NW warps>: <1,16> <1,20> <2,16> <2,20> <2,24> <3,16>), ties, more picks than points, more picks than the output
    ring holds."""
    rng = np.random.default_rng(N * 11 + npoint)
    xyz = scene(rng, B, N, quant, kind)
    want = C.furthest_point_sample(xyz, npoint)
    x = cu(xyz)
    grid = ops.build_scene_grid(x)
    for variant in (0, -2):           # 0: the instance with the fewest box-test rounds (16 / 20 / 24 warps); -2: 16 warps only
        got = ops.furthest_point_sample(x, npoint, grid, "throughput", False, variant)
        assert np.array_equal(got.cpu().numpy(), want), f"variant {variant}"


def test_throughput_fps_all_points_equal(ops):
    p = np.ones((2, 300, 3), np.float32)
    x = cu(p)
    assert ops.furthest_point_sample(x, 5, ops.build_scene_grid(x), "throughput").cpu().tolist() == [[0] * 5] * 2
