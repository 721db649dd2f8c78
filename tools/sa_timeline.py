"""CTA-0 timeline of the specialised SA kernel (needs the -DSAD_MLP_PROFILE build:
    python tools/build_variant.py prof -DSAD_MLP_PROFILE
    SAD_B200_LIB=3dsad-main_b200/lib/libsad_prof.so python tools/sa_timeline.py sa4 [first] [count])"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200  # noqa
from sad_b200 import mlp as M, _lib

lib = _lib.load()
dump = lib.sad_sa_profile_dump
dump.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
ROLES = ["epi0", "gath", "mmaA", "prod", "epi4", "mmaB"]


def layers(ch):
    g = torch.Generator().manual_seed(0)
    return M.prepare_layers([((torch.randn(co, ci, generator=g) / ci ** 0.5).cuda(), torch.zeros(co).cuda())
                             for ci, co in zip(ch[:-1], ch[1:])])


dbg = torch.zeros(64, dtype=torch.int64).pin_memory()
lib.sad_sa_debug_buffer.argtypes = [ctypes.c_void_p]
lib.sad_sa_debug_buffer(dbg.data_ptr())


def report_dbg():
    n = int(dbg[63]) & 0xFFFFFFFF
    print(f"barrier timeouts reported: {n}; misc base offsets: see Misc layout")
    for i in range(min(n, 15)):
        a, b = int(dbg[4 * i]), int(dbg[4 * i + 1])
        print(f"  block {a >> 32} thread {a & 0xFFFFFFFF} (warp {(a & 0xFFFFFFFF) >> 5}) bar smem addr {(b >> 32) & 0xFFFFFFFF:#x} parity {b & 0xFFFFFFFF}")


def timeline(tag, fn, first=0, count=120):
    try:
        fn(); torch.cuda.synchronize()
    except Exception as e:
        print("FAILED:", str(e).splitlines()[0])
        report_dbg()
        raise
    log = np.zeros((6, 8192), dtype=np.int64); n = np.zeros(6, dtype=np.int32)
    dump(log.ctypes.data, n.ctypes.data)
    fn(); torch.cuda.synchronize()
    dump(log.ctypes.data, n.ctypes.data)
    ev = []
    for role, name in enumerate(ROLES):
        for i in range(n[role]):
            ev.append((int(log[role, 2 * i + 1]), name, int(log[role, 2 * i])))
    ev.sort()
    t0 = ev[0][0]
    print(f"== {tag}: {len(ev)} events, CTA-0 span {(ev[-1][0] - t0) / 1.9e3:.1f} us (at 1.9 GHz)")
    for (t, name, e) in ev[first:first + count]:
        print(f"  {(t - t0):9d} cyc  {name} {e}")


SHAPES = {"sa1": (40000, 2048, 64, 1, [64, 64, 128]), "sa2": (2048, 1024, 32, 128, [128, 128, 256]),
          "sa3": (1024, 512, 16, 256, [128, 128, 256]), "sa4": (512, 256, 16, 256, [128, 128, 256])}
name = sys.argv[1] if len(sys.argv) > 1 else "sa4"
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 120
if len(sys.argv) > 4:
    M.FAST_SA[0] = sys.argv[4]
N, P, S, C, hid = SHAPES[name]
B, dev = 8, "cuda"
xyz = torch.rand(B, N, 3, device=dev)
new_xyz = torch.rand(B, P, 3, device=dev)
idx = torch.randint(0, N, (B, P, S), device=dev, dtype=torch.int32)
feat = torch.randn(B, C, N, device=dev)
if C >= 64:
    feat._sad_cl = M.to_cl_bf16(feat)
m = layers([C + 3] + hid)
timeline(f"{name} B={B}", lambda: M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.3, m), first, count)
