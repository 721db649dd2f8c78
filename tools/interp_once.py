"""One three_interpolate forward at the FP2 shape (for ncu):  python tools/interp_once.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200 as S  # noqa

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
C, m, n = 256, 512, 1024
g = torch.Generator(device="cuda").manual_seed(0)
f = torch.randn(B, C, m, device="cuda", generator=g)
idx = torch.randint(0, m, (B, n, 3), device="cuda", dtype=torch.int32, generator=g)
w = torch.rand(B, n, 3, device="cuda", generator=g)
for _ in range(3):
    out = S.three_interpolate(f, idx, w)
torch.cuda.synchronize()
print(float(out.sum()))
