"""Scene-level data parallelism: one process per GPU, scenes are
independent, so the inference path shards contiguous blocks of scenes across ranks with NO
data-path collective.  The only communication is bookkeeping: a barrier to align timed
windows and a MAX / SUM reduction of scalars.  Works on NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(num_scenes: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of scenes owned by `rank`; blocks differ by at most one scene."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(num_scenes, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def weak_scaling_first_scene(rank: int, scenes_per_rank: int, input_sets: int = 1, set_index: int = 0) -> int:
    """Seed offset of the first scene of (`rank`, rotating input set) under weak scaling: every
    rank gets its own distinct scenes."""
    return (rank * input_sets + set_index) * scenes_per_rank


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def reduce_scalars(values, op: str = "max", device=None):
    """All-reduce a list of python floats (MAX for times, SUM for counts) -> list of floats."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    world, _ = world_info()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return [float(x) for x in t]


def aggregate_rate(units_this_rank: float, seconds_this_rank: float, device=None) -> float:
    """Whole-job throughput: units processed by ALL ranks / MAX over ranks of the time."""
    (total_units,) = reduce_scalars([units_this_rank], "sum", device)
    (max_s,) = reduce_scalars([seconds_this_rank], "max", device)
    return total_units / max_s


def barrier():
    world, _ = world_info()
    if world > 1:
        dist.barrier()
