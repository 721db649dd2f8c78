"""Smallest program that runs the benchmarked step (B=8 x 40k hot path) -- the command
profiled under ncu (launch list / --set full captures).  Not a benchmark: prints nothing
but a checksum.

    python tools/profile_step.py [--warmup 2] [--steps 1] [--B 8] [--N 40000]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200  # noqa: E402,F401
from sad_b200.config import LAYER_CFG, make_params  # noqa: E402
from sad_b200.modules import SADHotPath  # noqa: E402
from sad_b200.scenes import make_scenes, make_sizes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--N", type=int, default=40000)
    ap.add_argument("--calls", action="store_true", help="print CUDA-event time of every C-ABI call (serial, one stream)")
    ap.add_argument("--fps-policy", default="throughput", choices=["throughput", "latency"],
                    help="which scene-grid FPS kernel runs (bench.py's pipelined executor captures `throughput`)")
    a = ap.parse_args()
    dev = "cuda:0"
    from sad_b200 import modules as _modules
    _modules.FPS_POLICY[0] = a.fps_policy
    model = SADHotPath(1).load_params(make_params(0)).to(dev).eval()
    xyz, feat = make_scenes(a.B, a.N, "surface")
    size = make_sizes(a.B, LAYER_CFG["agg"][0])
    x, f, s = (torch.from_numpy(t).to(dev) for t in (xyz, feat, size))
    with torch.no_grad():
        for _ in range(a.warmup + a.steps):
            end = model(x, f, s)
    torch.cuda.synchronize()
    if a.calls:
        from sad_b200 import _lib
        model.backbone.overlap_geometry = False
        acc = {}
        for rep in range(5):
            with _lib.CallProfiler() as prof, torch.no_grad():
                model(x, f, s)
            for i, (name, args, ms) in enumerate(prof.rows()):
                acc.setdefault((i, name, tuple(v for v in args[:4] if isinstance(v, int))), []).append(ms)
        tot = 0.0
        for (i, name, dims), v in acc.items():
            v.sort()
            tot += v[len(v) // 2]
            print(f"{i:3d} {name:34s} {str(dims):28s} {1e3 * v[len(v) // 2]:9.1f} us")
        print(f"sum of C-ABI calls: {1e3 * tot:.1f} us")
    print("checksum", float(end["cluster_features"].sum()))


if __name__ == "__main__":
    main()
