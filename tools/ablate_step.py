"""Tools only: how much of the pipelined step is the 40k-point FPS?  Runs bench.py's product arm with the scene-grid
FPS replaced by a strided index list (WRONG results, timing only) so the step time without the sampling chain can be
compared with the real one.   python tools/ablate_step.py [nofps|nomlp] <bench.py args>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sad_b200  # noqa
from sad_b200 import ops, modules
import bench

what = sys.argv[1]
sys.argv = [sys.argv[0]] + sys.argv[2:]
if what == "nofps":
    real = ops.furthest_point_sample

    def fake(xyz, npoint, grid=None, policy="latency", prefix_ordered=False):
        if grid is None:
            return real(xyz, npoint, grid, policy, prefix_ordered)
        B, N, _ = xyz.shape
        return (torch.arange(npoint, device=xyz.device, dtype=torch.int32) * (N // npoint)).unsqueeze(0).repeat(B, 1).contiguous()

    ops.furthest_point_sample = fake
sys.exit(bench.main())
