"""Per-phase cycle breakdown of the FPS kernel (block 0) via sad_fps_set_debug_buffer."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200 as S
from sad_b200 import _lib
from sad_b200.scenes import make_scenes
lib = _lib.load()
dbg = torch.zeros(8, dtype=torch.int64, device="cuda:0")
names = ["local pass", "warp reduce", "cta barrier", "exchange", "selection"]
for (N, npnt, cs) in [(40000, 2048, 0), (40000, 2048, 8), (2048, 1024, 0), (1024, 512, 0), (1024, 256, 0)]:
    xyz = torch.from_numpy(make_scenes(8, max(N, 2048), "surface")[0][:, :N].copy()).cuda()
    lib.sad_fps_force_cluster_size(cs)
    lib.sad_fps_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
    S.furthest_point_sample(xyz, npnt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    S.furthest_point_sample(xyz, npnt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    lib.sad_fps_set_debug_buffer(None)
    lib.sad_fps_force_cluster_size(0)
    d = dbg.cpu().tolist()
    rounds = max(1, d[5])
    tot = sum(d[:5])
    print(f"  kernel {ms:.3f} ms, block-0 lifetime {d[6]} cycles => {d[6] / ms / 1e6:.2f} GHz-equivalent")
    print(f"N={N} npoint={npnt} cs={cs or 'auto'}: rounds={rounds} picks/round={(npnt - 1) / rounds:.2f} "
          f"cycles/round={tot / rounds:.0f} cycles/pick={tot / (npnt - 1):.0f}  " +
          "  ".join(f"{n}={v / rounds:.0f}" for n, v in zip(names, d[:5])))
