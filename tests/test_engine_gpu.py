"""Pipelined executor (engine.py): CUDA-graph replays return exactly what the eager forward returns,
in submission order, for host and device callers."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def setup():
    import sad_b200  # noqa: F401
    from sad_b200.config import LAYER_CFG, make_params
    from sad_b200.engine import PipelinedHotPath
    from sad_b200.modules import SADHotPath
    from sad_b200.scenes import make_scenes, make_sizes
    model = SADHotPath(1).load_params(make_params(0)).to(DEV).eval()
    B, N = 2, 9000
    eng = PipelinedHotPath(model, B, N, slots=3, device=torch.device(DEV))
    batches = []
    for k in range(5):
        xyz, feat = make_scenes(B, N, "surface", first_scene=10 * k)
        size = make_sizes(B, LAYER_CFG["agg"][0], first_scene=10 * k)
        batches.append(tuple(torch.from_numpy(a).pin_memory() for a in (xyz, feat, size)))
    return model, eng, batches


def test_graph_replay_equals_eager_and_keeps_order(setup):
    model, eng, batches = setup
    want = []
    with torch.no_grad():
        for h in batches:
            end = model(*(t.to(DEV) for t in h))
            want.append((end["cluster_xyz"].cpu().numpy(), end["cluster_features"].cpu().numpy(),
                         end["cluster_inds"].cpu().numpy(), end["sa1_inds"].cpu().numpy()))
    tickets, got = [], []
    for h in batches:                       # more batches than slots: back-pressure + slot reuse
        tickets.append(eng.submit_host(*h))
        if len(tickets) == eng.slots:
            t = tickets.pop(0)
            cx, cf = eng.result(t)
            got.append((cx.numpy().copy(), cf.numpy().copy(), eng.outputs(t)["cluster_inds"].cpu().numpy(),
                        eng.outputs(t)["sa1_inds"].cpu().numpy()))
    for t in tickets:
        cx, cf = eng.result(t)
        got.append((cx.numpy().copy(), cf.numpy().copy(), eng.outputs(t)["cluster_inds"].cpu().numpy(),
                    eng.outputs(t)["sa1_inds"].cpu().numpy()))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert np.array_equal(g[3], w[3]) and np.array_equal(g[2], w[2])       # sampling indices: bit-exact
        assert np.array_equal(g[0], w[0])                                       # cluster centres: bit-exact
        assert np.array_equal(g[1], w[1])                                       # same kernels, same order: identical


def test_device_submission_and_launch_count(setup):
    model, eng, batches = setup
    dev_in = tuple(t.to(DEV) for t in batches[0])
    t = eng.submit_device(*dev_in, to_host=True)
    cx, cf = eng.result(t)
    with torch.no_grad():
        end = model(*dev_in)
    assert np.array_equal(cf.numpy(), end["cluster_features"].cpu().numpy())
    assert eng.launches_per_batch >= 20      # kernels of libsad_b200 captured per batch


def test_resident_submission_and_forked_pytorch_path(setup):
    """submit_resident (zero-copy: the batch is written straight into the slot's input buffers) and the executor's other
    configuration (forked graph, PyTorch-call submission) return exactly what the default one does."""
    from sad_b200.engine import PipelinedHotPath
    model, eng, batches = setup
    dev_in = tuple(t.to(DEV) for t in batches[1])
    t = eng.submit_device(*dev_in, to_host=True)
    cx, cf = eng.result(t)
    want_x, want_f = cx.numpy().copy(), cf.numpy().copy()
    for dst, src in zip(eng.slot_inputs(eng.next_slot), dev_in):
        dst.copy_(src)
    ev = torch.cuda.Event()
    ev.record()
    t = eng.submit_resident(after=ev, to_host=True)
    cx, cf = eng.result(t)
    assert np.array_equal(cx.numpy(), want_x) and np.array_equal(cf.numpy(), want_f)
    other = PipelinedHotPath(model, eng.batch, eng.n_points, slots=2, device=torch.device(DEV), native_submit=False,
                             linear_graph=False)
    assert not other.native_submit and not other.linear_graph and eng.native_submit and eng.linear_graph
    for submit in (lambda: other.submit_host(*batches[1]), lambda: other.submit_device(*dev_in, to_host=True)):
        cx, cf = other.result(submit())
        assert np.array_equal(cx.numpy(), want_x) and np.array_equal(cf.numpy(), want_f)
    for dst, src in zip(other.slot_inputs(other.next_slot), dev_in):
        dst.copy_(src)
    torch.cuda.synchronize()
    cx, cf = other.result(other.submit_resident(to_host=True))
    assert np.array_equal(cx.numpy(), want_x) and np.array_equal(cf.numpy(), want_f)


def _close_elementwise(got, want, tol, rel, floor_frac=0.25, what=""):
    """Norm-wise bar (max |err| <= tol * max |want|) plus an element-wise relative bar above a magnitude floor."""
    scale = max(1e-6, float(np.abs(want).max()))
    err = np.abs(got - want)
    assert float(err.max()) <= tol * scale, f"{what}: max abs err {float(err.max()):.4g} vs scale {scale:.4g} (tol {tol})"
    m = np.abs(want) > floor_frac * scale
    if m.any():
        r = float((err[m] / np.abs(want)[m]).max())
        assert r <= rel, f"{what}: max element-wise relative err {r:.4g} above {floor_frac} of scale (tol {rel})"


def test_benchmarked_configuration_matches_oracle():
    """the EXACT configuration bench.py times -- 8 x 40k surface scenes per batch through
    PipelinedHotPath(fps_policy="throughput", mlp_tiles_per_cta=6), i.e. the one-SM-per-scene FPS, the specialised SA
    kernels on narrow grids, the prefix-ordered sampling shortcut, CUDA-graph replay -- against the oracle end to end:
    every index tensor bit for bit, features within the bf16 bar: norm-wise 2e-2 of the tensor's scale, plus
    element-wise 6e-2 relative for every element above a quarter of that scale (smaller elements are sums that
    cancel: their absolute error is bounded by the norm-wise bar, their relative error is not meaningful)."""
    import sad_b200  # noqa: F401
    from oracle import sad_oracle as O, c_port as C
    from sad_b200.config import LAYER_CFG, make_params
    from sad_b200.engine import PipelinedHotPath
    from sad_b200.modules import SADHotPath
    from sad_b200.scenes import make_scenes, make_sizes
    C.build()
    params = make_params(0)
    model = SADHotPath(1).load_params(params).to(DEV).eval()
    B, N = 8, 40000
    eng = PipelinedHotPath(model, B, N, slots=4, device=torch.device(DEV), fps_policy="throughput", mlp_tiles_per_cta=6)
    hosts, tickets = [], []
    for k in range(2):
        xyz, feat = make_scenes(B, N, "surface", first_scene=100 + B * k)
        size = make_sizes(B, LAYER_CFG["agg"][0], first_scene=100 + B * k)
        hosts.append((xyz, feat, size))
        tickets.append(eng.submit_host(*(torch.from_numpy(a).pin_memory() for a in (xyz, feat, size))))
    for (xyz, feat, size), t in zip(hosts, tickets):
        cx, cf = eng.result(t)
        end = eng.outputs(t)
        want = O.backbone_forward(xyz, feat, params, LAYER_CFG, False, C)
        for key in ("sa1_inds", "sa2_inds", "sa3_inds", "sa4_inds"):
            assert np.array_equal(end[key].cpu().numpy(), want[key]), f"{key} differs from the oracle"
        for key in ("sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz"):
            assert np.array_equal(end[key].cpu().numpy(), want[key]), f"{key} differs from the oracle"
        for key in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
            _close_elementwise(end[key].cpu().numpy(), want[key], 2e-2, 6e-2, what=key)
        # voting + clustering head: the votes are bf16-feature dependent, so the GPU's own votes feed the oracle's
        # head and its indices / centres must then match bit for bit (same protocol as tests/test_modules_gpu.py)
        wv_xyz, wv_feat = O.voting_module(want["fp2_xyz"], end["fp2_features"].cpu().numpy(), params["vote"])
        _close_elementwise(end["vote_xyz"].cpu().numpy(), wv_xyz, 2e-2, 6e-2, what="vote_xyz")
        _close_elementwise(end["vote_features"].cpu().numpy(), wv_feat, 2e-2, 6e-2, what="vote_features")
        npoint, _, nsample = LAYER_CFG["agg"]
        cxyz, cfeat, cinds, rt = O.vote_aggregation(
            end["vote_xyz"].cpu().numpy(), end["vote_features"].cpu().numpy(), size, npoint, nsample, params["agg"],
            alpha=LAYER_CFG["alpha"], r_min=LAYER_CFG["r_min"], r_max=LAYER_CFG["r_max"], impl=C)
        assert np.array_equal(end["cluster_inds"].cpu().numpy(), cinds), "cluster_inds differ from the oracle"
        assert np.array_equal(cx.numpy(), cxyz), "cluster centres differ from the oracle"
        assert np.array_equal(end["cluster_radius"].cpu().numpy(), rt), "per-cluster radii differ from the oracle"
        _close_elementwise(cf.numpy(), cfeat, 2e-2, 6e-2, what="cluster_features")
