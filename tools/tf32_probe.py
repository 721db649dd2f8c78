"""One launch of the tf32 fused stage per benchmark shape (SAD_B200_LIB=...libsad_tfprof.so prints CTA-0 wait times)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200  # noqa
from sad_b200 import mlp as M


def layers(ch):
    g = torch.Generator().manual_seed(0)
    return M.prepare_layers([((torch.randn(co, ci, generator=g) / ci ** 0.5).cuda(), 0.1 * torch.randn(co, generator=g).cuda())
                             for ci, co in zip(ch[:-1], ch[1:])], dtype="tf32")


def timed(fn, name):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    print("==", name, flush=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    print("   ", name, round(a.elapsed_time(b) * 1e3, 1), "us", flush=True)


B = 8
which = sys.argv[1:] or ["sa1", "sa2", "fp1"]
g = torch.Generator(device="cuda").manual_seed(0)
if "sa1" in which:
    N, P, S = 40000, 2048, 64
    xyz = torch.rand(B, N, 3, device="cuda") * 6
    feat = torch.randn(B, 1, N, device="cuda")
    idx = torch.randint(0, N, (B, P, S), device="cuda", dtype=torch.int32)
    new_xyz = xyz[:, :P].contiguous()
    mlp = layers([4, 64, 64, 128])
    timed(lambda: M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.2, mlp), "sa1")
if "sa2" in which:
    N, P, S = 2048, 1024, 32
    xyz = torch.rand(B, N, 3, device="cuda") * 6
    feat = torch.randn(B, 128, N, device="cuda")
    idx = torch.randint(0, N, (B, P, S), device="cuda", dtype=torch.int32)
    new_xyz = xyz[:, :P].contiguous()
    mlp = layers([131, 128, 128, 256])
    timed(lambda: M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.4, mlp), "sa2")
if "fp1" in which:
    n, m = 512, 256
    kf = torch.randn(B, 256, m, device="cuda")
    sf = torch.randn(B, 256, n, device="cuda")
    idx = torch.randint(0, m, (B, n, 3), device="cuda", dtype=torch.int32)
    w = torch.rand(B, n, 3, device="cuda")
    w = (w / w.sum(-1, keepdim=True)).contiguous()
    mlp = layers([512, 256, 256])
    timed(lambda: M.fp_interp_mlp(kf, sf, idx, w, mlp), "fp1")
