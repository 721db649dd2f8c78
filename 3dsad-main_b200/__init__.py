"""sad_b200 -- B200-native set-abstraction / size-adaptive clustering hot path.

Python host side of libsad_b200.so (hand-written CUDA for sm_100a behind a plain C
ABI, include/sad_ops.h).  It mirrors the PointNet++-lineage operator surface that
BASELINE.json's north_star names (the mounted reference, /root/reference/README.md:1-2,
contains no code to cite): furthest_point_sample, gather_operation, ball_query,
ball_query_adaptive, grouping_operation, three_nn, three_interpolate as
torch.autograd.Function callables, plus the SA / FP / vote-aggregation modules built
from them.  There is no CPU fallback: importing the ops without the built extension,
or calling them on CPU tensors, raises.
"""
from . import _lib  # noqa: F401
from .ops import (  # noqa: F401
    furthest_point_sample, gather_operation, ball_query, ball_query_adaptive,
    grouping_operation, three_nn, three_interpolate, size_to_radius, set_deterministic,
    FurthestPointSampling, GatherOperation, BallQuery, BallQueryAdaptive,
    GroupingOperation, ThreeNN, ThreeInterpolate,
)

__version__ = "0.1.0"
