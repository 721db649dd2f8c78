"""Does a batch of 16-CTA FPS clusters run concurrently?  Kernel time vs number of scenes."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200 as S
from sad_b200 import _lib
from sad_b200.scenes import make_scenes
lib = _lib.load()
full = torch.from_numpy(make_scenes(16, 40000, "surface")[0]).cuda()
for cs in (16, 14, 12):
    for B in (1, 6, 7, 8, 9, 10):
        if B * cs > 400:
            continue
        x = full[:B].contiguous()
        lib.sad_fps_force_cluster_size(cs)
        for _ in range(2):
            S.furthest_point_sample(x, 2048)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        S.furthest_point_sample(x, 2048)
        e1.record()
        torch.cuda.synchronize()
        print(f"cs={cs:2d} B={B:2d}  {e0.elapsed_time(e1):.3f} ms", flush=True)
lib.sad_fps_force_cluster_size(0)
