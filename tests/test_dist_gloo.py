"""N>1 host logic on CPU: world_size-2 gloo processes exercise the scene sharding and the
max-over-ranks / sum-over-ranks bookkeeping bench.py uses; the
data path itself needs no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sad_b200  # noqa: F401
        from sad_b200 import dist as D
        from sad_b200.scenes import make_scenes
        from oracle import c_port as C

        lo, hi = D.shard_range(5, world, rank)
        # each rank runs the (CPU-oracle) FPS on its own shard; rank results are independent
        xyz, _ = make_scenes(hi - lo, 600, "uniform", first_scene=lo)
        inds = C.furthest_point_sample(xyz, 16)
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, inds))
        rate = D.aggregate_rate(units_this_rank=float(hi - lo), seconds_this_rank=0.5 + rank)
        mx = D.reduce_scalars([float(rank + 1), 10.0 - rank], "max")
        first = D.weak_scaling_first_scene(rank, 8, input_sets=4, set_index=2)
        D.barrier()
        q.put((rank, gathered, rate, mx, first))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    import sad_b200  # noqa: F401
    from sad_b200.dist import shard_range
    for n in (0, 1, 5, 8, 17, 64):
        for w in (1, 2, 3, 8):
            blocks = [shard_range(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


@pytest.mark.timeout(180)
def test_two_rank_gloo_sharding_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    from sad_b200.scenes import make_scenes
    from oracle import c_port as C
    xyz, _ = make_scenes(5, 600, "uniform")
    want = C.furthest_point_sample(xyz, 16)
    for rank, gathered, rate, mx, first in results:
        got = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
        assert np.array_equal(got, want)                       # sharded == unsharded, scene for scene
        assert abs(rate - 5.0 / 1.5) < 1e-12                   # all scenes / slowest rank
        assert mx == [2.0, 10.0]
        assert first == (rank * 4 + 2) * 8


class _StubEngine:
    """CPU stand-in for engine.PipelinedHotPath: same submit_host / result / slots surface, a per-scene function of the
    inputs as the 'result' (so the test can tell exactly which scenes went where, padding included)."""
    slots = 2

    def __init__(self):
        self._pending = {}
        self._n = 0

    def submit_host(self, xyz, feat, size):
        t = self._n
        self._n += 1
        self._pending[t] = (xyz.sum(dim=(1, 2)).reshape(-1, 1, 1) + torch.zeros(xyz.shape[0], 4, 3),
                            feat.mean(dim=(1, 2)).reshape(-1, 1, 1) + size.sum(dim=(1, 2)).reshape(-1, 1, 1) + torch.zeros(xyz.shape[0], 2, 4))
        return t

    def result(self, t):
        return self._pending.pop(t)


def _sharded_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sad_b200  # noqa: F401
        from sad_b200.engine import ShardedHotPath
        g = torch.Generator().manual_seed(5)
        S = 11                                         # not a multiple of world * batch: ragged shards, padded last batches
        xyz, feat, size = torch.rand(S, 50, 3, generator=g), torch.rand(S, 1, 50, generator=g), torch.rand(S, 4, 3, generator=g)
        shp = ShardedHotPath(batch=3, engine=_StubEngine())
        cx, cf = shp.run(xyz, feat, size)
        q.put((rank, shp.my_range(S), cx.numpy(), cf.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_hot_path_two_ranks_equal_one_process():
    """engine.ShardedHotPath (the product's multi-GPU entry point) on 2 gloo ranks with a stub engine: rank 0 receives
    every scene's result in scene order, equal to one process running all scenes."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {r[0]: r for r in [q.get(timeout=150) for _ in range(world)]}
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    import sad_b200  # noqa: F401
    from sad_b200.engine import ShardedHotPath
    g = torch.Generator().manual_seed(5)
    S = 11
    xyz, feat, size = torch.rand(S, 50, 3, generator=g), torch.rand(S, 1, 50, generator=g), torch.rand(S, 4, 3, generator=g)
    one = ShardedHotPath(batch=3, engine=_StubEngine())          # no process group: a single process owns every scene
    wx, wf = one.run(xyz, feat, size)
    assert wx.shape[0] == S
    assert np.array_equal(results[0][2], wx.numpy()) and np.array_equal(results[0][3], wf.numpy())
    lo, hi = results[1][1]
    assert (lo, hi) == (6, 11)
    assert np.array_equal(results[1][2], wx.numpy()[lo:hi])      # the other rank keeps its own block
