"""Oracle self-tests: hand-derived known-answer vectors (tier 1)
and NumPy-oracle == C-port bit-equality.  CPU only.

The reference mount is README-only, so these hand-checkable cases (worked out in
the comments) are what pins the oracle."""
import numpy as np
import pytest

from oracle import sad_oracle as O
from oracle import c_port as C


def line(n, step=1.0):
    p = np.zeros((1, n, 3), dtype=np.float32)
    p[0, :, 0] = np.arange(n, dtype=np.float32) * np.float32(step)
    return p


IMPLS = [pytest.param(O, id="numpy"), pytest.param(C, id="c_port")]


# ------------------------------------------------------------------ FPS (a1)
@pytest.mark.parametrize("impl", IMPLS)
def test_fps_collinear_known_answer(impl):
    # x = 0..9.  start 0 -> farthest 9.  mind = min(d0, d9): x=4 -> min(16,25)=16 and
    # x=5 -> min(25,16)=16 tie -> lowest index 4.  With {0,9,4}: mind = [0,1,4,1,0,1,4,4,1,0]
    # -> first max at 2.  With {0,9,4,2}: x=6 -> 4, x=7 -> 4 -> 6.  Then x=7: min(49,4,9,25,1)=1;
    # remaining mind are all 1 -> lowest index 1.
    got = impl.furthest_point_sample(line(10), 6)
    assert got.dtype == np.int32
    assert got.tolist() == [[0, 9, 4, 2, 6, 1]]


@pytest.mark.parametrize("impl", IMPLS)
def test_fps_duplicates_pick_lowest_index(impl):
    p = np.ones((2, 7, 3), dtype=np.float32)
    assert impl.furthest_point_sample(p, 4).tolist() == [[0, 0, 0, 0]] * 2


@pytest.mark.parametrize("impl", IMPLS)
def test_fps_npoint_one_and_npoint_gt_N(impl):
    assert impl.furthest_point_sample(line(5), 1).tolist() == [[0]]
    # more samples than points: after all 3 distinct points are taken every mind is 0
    # -> argmax of an all-zero array = 0 for ever after.
    assert impl.furthest_point_sample(line(3), 6).tolist() == [[0, 2, 1, 0, 0, 0]]


def test_fps_symmetric_tie():
    # square corners + centre first: all four corners are equidistant from the centre.
    p = np.array([[[0, 0, 0], [1, 1, 0], [-1, 1, 0], [1, -1, 0], [-1, -1, 0]]], dtype=np.float32)
    # step1: four-way tie (d2=2) -> index 1.  step2: from {0,1}: idx2 -> min(2,4)=2, idx3 -> 2,
    # idx4 -> min(2,8)=2 -> lowest = 2.  step3: idx3: min(2,4,8)=2 ; idx4: min(2,8,4)=2 -> 3.
    assert O.furthest_point_sample(p, 5).tolist() == [[0, 1, 2, 3, 4]]
    assert C.furthest_point_sample(p, 5).tolist() == [[0, 1, 2, 3, 4]]


# ----------------------------------------------------------- ball query (a3/a4)
@pytest.mark.parametrize("impl", IMPLS)
def test_ball_query_known_answers(impl):
    xyz = line(5, 0.5)                                      # x = 0, .5, 1, 1.5, 2
    q = np.array([[[0, 0, 0], [10, 0, 0], [1, 0, 0]]], dtype=np.float32)
    got = impl.ball_query(1.0, 4, xyz, q)
    # q0: d2 = 0, .25, 1.0 (NOT < 1.0: strict), ... -> hits {0,1}, padded with first hit 0.
    # q1: empty ball -> zeros.   q2 (x=1): d2 = 1, .25, 0, .25, 1 -> hits {1,2,3}, pad 1.
    assert got.tolist() == [[[0, 1, 0, 0], [0, 0, 0, 0], [1, 2, 3, 1]]]
    # more hits than nsample: first nsample in ascending index order
    assert impl.ball_query(5.0, 2, xyz, q)[0, 0].tolist() == [0, 1]
    assert impl.ball_query(5.0, 1, xyz, q)[0, 2].tolist() == [0]


@pytest.mark.parametrize("impl", IMPLS)
def test_ball_query_adaptive_per_query_radius(impl):
    xyz = line(5, 0.5)
    q = np.array([[[1, 0, 0], [1, 0, 0], [1, 0, 0]]], dtype=np.float32)
    r = np.array([[0.25, 0.75, 1.25]], dtype=np.float32)
    got = impl.ball_query_adaptive(r, 5, xyz, q)
    # r=.25: only x=1 (d2=0; .25 is not < .0625).  r=.75: {1,2,3}.  r=1.25: all five.
    assert got.tolist() == [[[2, 2, 2, 2, 2], [1, 2, 3, 1, 1], [0, 1, 2, 3, 4]]]
    # adaptive with a constant radius == plain ball query
    rc = np.full((1, 3), 0.75, dtype=np.float32)
    assert np.array_equal(impl.ball_query_adaptive(rc, 5, xyz, q), impl.ball_query(0.75, 5, xyz, q))


def test_size_to_radius():
    s = np.array([[[0.6, 0.8, 0.0], [0.02, 0.0, 0.0], [10, 10, 10]]], dtype=np.float32)
    r = O.size_to_radius(s, alpha=1.0, r_min=0.1, r_max=1.2)
    assert r.dtype == np.float32
    np.testing.assert_allclose(r, [[0.5, 0.1, 1.2]], rtol=1e-6)


# ----------------------------------------------------------------- three_nn (a8)
@pytest.mark.parametrize("impl", IMPLS)
def test_three_nn_known_answer_and_ties(impl):
    known = line(4)                                           # x = 0,1,2,3
    unknown = np.array([[[1.5, 0, 0], [0, 0, 0]]], dtype=np.float32)
    dist, idx = impl.three_nn(unknown, known)
    # 1.5: d2 = 2.25,.25,.25,2.25 -> (1, 2, 0): both ties go to the lower index.
    assert idx.tolist() == [[[1, 2, 0], [0, 1, 2]]]
    np.testing.assert_array_equal(dist, np.array([[[0.5, 0.5, 1.5], [0, 1, 2]]], dtype=np.float32))
    with pytest.raises(ValueError):
        impl.three_nn(unknown, line(2))


def test_interpolation_weights_sum_to_one():
    rng = np.random.default_rng(0)
    d = rng.random((2, 50, 3), dtype=np.float32)
    d[0, 0] = 0.0                                             # coincident points stay finite
    w = O.interpolation_weights(d)
    assert np.isfinite(w).all()
    np.testing.assert_allclose(w.sum(-1), 1.0, rtol=1e-6)


# ------------------------------------------------------ gather / group / interp
def test_gather_group_interpolate_known_answers():
    f = np.arange(12, dtype=np.float32).reshape(1, 2, 6)      # c0: 0..5, c1: 6..11
    idx = np.array([[5, 0, 3]], dtype=np.int32)
    assert O.gather_operation(f, idx).tolist() == [[[5, 0, 3], [11, 6, 9]]]
    assert C.gather_operation(f, idx).tolist() == [[[5, 0, 3], [11, 6, 9]]]
    gidx = np.array([[[1, 1], [4, 2]]], dtype=np.int32)
    want = [[[[1, 1], [4, 2]], [[7, 7], [10, 8]]]]
    assert O.grouping_operation(f, gidx).tolist() == want
    assert C.grouping_operation(f, gidx).tolist() == want
    # backward: index 1 receives two contributions
    g = O.grouping_operation_grad(np.ones((1, 2, 2, 2), np.float32), gidx, 6)
    assert g[0, 0].tolist() == [0, 2, 1, 0, 1, 0]
    gg = O.gather_operation_grad(np.ones((1, 2, 3), np.float32), np.array([[1, 1, 2]], np.int32), 4)
    assert gg[0, 1].tolist() == [0, 2, 1, 0]
    # interpolate
    ii = np.array([[[0, 1, 2], [3, 3, 3]]], dtype=np.int32)
    w = np.array([[[0.5, 0.25, 0.25], [1, 0, 0]]], dtype=np.float32)
    out = O.three_interpolate(f, ii, w)
    assert out.tolist() == [[[0.75, 3.0], [6.75, 9.0]]]
    assert C.three_interpolate(f, ii, w).tolist() == out.tolist()
    gi = O.three_interpolate_grad(np.ones((1, 2, 2), np.float32), ii, w, 6)
    assert gi[0, 0].tolist() == [0.5, 0.25, 0.25, 1.0, 0, 0]


def test_shared_mlp_known_answer():
    # one layer, identity-ish weights: y = relu(W x + b), max over samples
    x = np.array([[[[1, -2, 3]], [[-1, -1, -1]]]], dtype=np.float32)       # (1,2,1,3)
    W = np.array([[1, 0], [0, 1], [1, 1]], dtype=np.float32)
    b = np.array([0, 0.5, 0], dtype=np.float32)
    y = O.shared_mlp(x, [(W, b)], pool=True)
    # rows: (1,-1)->[1,0,0]; (-2,-1)->[0,0,0]; (3,-1)->[3,0,2]  -> max = [3,0,2]
    assert y.tolist() == [[[3.0], [0.0], [2.0]]]
    y2 = O.shared_mlp(x, [(W, b)], pool=False, last_relu=False)
    assert y2[0, :, 0, 1].tolist() == [-2.0, -0.5, -3.0]


def test_bf16_round():
    a = np.array([1.0, 1.00390625, 1.01171875, -3.1415927], dtype=np.float32)
    r = O.bf16_round(a)
    # 1+2^-8 is exactly half-way between bf16 neighbours 1 and 1+2^-7: ties-to-even -> 1.0
    assert r[0] == 1.0 and r[1] == 1.0 and r[2] == np.float32(1.015625)
    assert abs(r[3] + 3.140625) < 1e-6


# ---------------------------------------------- NumPy oracle == C port, bit-exact
def _scene(rng, B, N, quant=None):
    p = rng.random((B, N, 3), dtype=np.float32) * np.float32(4.0) - np.float32(2.0)
    if quant:                                     # coarse lattice => many exact ties/duplicates
        p = np.round(p / np.float32(quant)) * np.float32(quant)
    return p.astype(np.float32)


@pytest.mark.parametrize("quant", [None, 0.25, 1.0])
def test_c_port_matches_numpy_bitexact(quant):
    rng = np.random.default_rng(7)
    xyz = _scene(rng, 3, 777, quant)
    inds = O.furthest_point_sample(xyz, 65)
    assert np.array_equal(inds, C.furthest_point_sample(xyz, 65))
    new_xyz = np.stack([xyz[b][inds[b]] for b in range(3)])
    for r, ns in [(0.3, 16), (0.75, 5), (5.0, 64)]:
        assert np.array_equal(O.ball_query(r, ns, xyz, new_xyz), C.ball_query(r, ns, xyz, new_xyz))
    rt = (rng.random((3, 65), dtype=np.float32) * np.float32(1.1) + np.float32(0.1)).astype(np.float32)
    assert np.array_equal(O.ball_query_adaptive(rt, 16, xyz, new_xyz),
                          C.ball_query_adaptive(rt, 16, xyz, new_xyz))
    d0, i0 = O.three_nn(xyz, new_xyz)
    d1, i1 = C.three_nn(xyz, new_xyz)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    f = rng.standard_normal((3, 9, 65)).astype(np.float32)
    w = O.interpolation_weights(d0)
    assert np.array_equal(O.three_interpolate(f, i0, w), C.three_interpolate(f, i0, w))
    idx = O.ball_query(0.75, 5, xyz, new_xyz)
    ff = rng.standard_normal((3, 9, 777)).astype(np.float32)
    assert np.array_equal(O.grouping_operation(ff, idx), C.grouping_operation(ff, idx))


def test_fps_nested_prefix_property():
    # FPS is prefix-nested: the first k picks do not depend on npoint.
    rng = np.random.default_rng(3)
    xyz = _scene(rng, 1, 400)
    a = O.furthest_point_sample(xyz, 64)
    b = O.furthest_point_sample(xyz, 17)
    assert np.array_equal(a[:, :17], b)


def test_sa_fp_modules_shapes_and_consistency():
    rng = np.random.default_rng(11)
    xyz = _scene(rng, 2, 300)
    feat = rng.standard_normal((2, 5, 300)).astype(np.float32)
    layers = [(rng.standard_normal((8, 8)).astype(np.float32) * 0.3, np.zeros(8, np.float32)),
              (rng.standard_normal((6, 8)).astype(np.float32) * 0.3, np.zeros(6, np.float32))]
    new_xyz, nf, inds = O.sa_module(xyz, feat, 32, 0.8, 8, layers)
    assert new_xyz.shape == (2, 32, 3) and nf.shape == (2, 6, 32) and inds.shape == (2, 32)
    assert (nf >= 0).all()
    fpl = [(rng.standard_normal((4, 11)).astype(np.float32), np.zeros(4, np.float32))]
    up = O.fp_module(xyz, new_xyz, feat, nf, fpl)
    assert up.shape == (2, 4, 300)
    # adaptive SA with a constant radius == plain SA
    rt = np.full((2, 32), 0.8, np.float32)
    _, nf2, _ = O.sa_module(xyz, feat, 32, None, 8, layers, radius_t=rt)
    np.testing.assert_array_equal(nf, nf2)
