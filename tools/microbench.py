"""Per-op CUDA-event timings at BASELINE config shapes (run under gpurun).

    python tools/microbench.py [--out gpurun_out/microbench.json] [--B 8] [--N 40000]

Not the bench contract (that is bench.py): this is the per-kernel view used to pick
what to optimise, with the algorithmic-bytes roofline of SURVEY.md section 8(d)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200 as S  # noqa: E402
from sad_b200 import _lib  # noqa: E402


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def timeit(fn, iters=20, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def surface_scene(g, B, N):
    """Room-like synthetic scene (SURVEY 8(d) distribution S, simplified): floor + walls + boxes."""
    pts = []
    for _ in range(B):
        u = torch.rand(N, 3, generator=g)
        which = torch.randint(0, 6, (N,), generator=g)
        p = torch.empty(N, 3)
        p[:, 0] = u[:, 0] * 6 - 3
        p[:, 1] = u[:, 1] * 6 - 3
        p[:, 2] = u[:, 2] * 3
        p[which == 0, 2] = 0
        p[which == 1, 0] = -3
        p[which == 2, 0] = 3
        p[which == 3, 1] = -3
        p[which == 4, 1] = 3
        p += torch.randn(N, 3, generator=g) * 0.005
        pts.append(p)
    return torch.stack(pts).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "microbench.json"))
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--N", type=int, default=40000)
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    dev = "cuda:0"
    B, N = args.B, args.N
    hbm, how = peaks()
    g = torch.Generator().manual_seed(1234)
    xyz = surface_scene(g, B, N).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {"B": B, "N": N, "hbm_peak_gbs": hbm, "peak": how, "rows": []}

    def row(name, ms, best, alg_bytes=None, note=""):
        r = {"op": name, "ms_median": round(ms, 4), "ms_best": round(best, 4), "note": note}
        if alg_bytes:
            r["alg_MB"] = round(alg_bytes / 1e6, 3)
            r["GBps"] = round(alg_bytes / best / 1e6, 1)
            r["hbm_frac"] = round(alg_bytes / best / 1e6 / hbm, 4)
        res["rows"].append(r)
        print(json.dumps(r), flush=True)

    # ---- FPS at every cluster size
    lib = _lib.load()
    for cs in (0, 4, 8, 16):
        try:
            ms, best = timeit(lambda: S.furthest_point_sample(xyz, 2048, None, "latency", False, cs), iters=10, warm=2)
            row(f"fps N={N}->2048 cs={cs or 'auto'}", ms, best, note=f"{2047 / best:.0f} iters/ms")
        except Exception as e:  # noqa: BLE001
            print("fps", cs, "failed:", e)
    inds = S.furthest_point_sample(xyz, 2048)
    new_xyz = S.gather_operation(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    for (n_in, n_out) in ((2048, 1024), (1024, 512), (512, 256)):
        x = new_xyz[:, :n_in].contiguous()
        ms, best = timeit(lambda: S.furthest_point_sample(x, n_out))
        row(f"fps N={n_in}->{n_out}", ms, best, note=f"{(n_out - 1) / best:.0f} iters/ms")

    # ---- ball query (SA1..SA4 shapes)
    cur = xyz
    layers = [(2048, 0.2, 64), (1024, 0.4, 32), (512, 0.8, 16), (256, 1.2, 16)]
    idxs = []
    for (npnt, r, ns) in layers:
        ii = S.furthest_point_sample(cur, npnt)
        q = S.gather_operation(cur.transpose(1, 2).contiguous(), ii).transpose(1, 2).contiguous()
        ms, best = timeit(lambda: S.ball_query(r, ns, cur, q), flush=flush)
        nin = cur.shape[1]
        row(f"ball_query N={nin} q={npnt} ns={ns}", ms, best, alg_bytes=B * (nin * 12 + npnt * 12 + npnt * ns * 4),
            note=f"{B * nin * npnt / best / 1e6:.1f} Gpair/s brute-equivalent")
        idxs.append((cur, q, S.ball_query(r, ns, cur, q)))
        cur = q

    # ---- grouping (channel-first surface) at SA1..SA4 feature widths
    for (src, q, idx), C in zip(idxs, (1, 128, 256, 256)):
        nin, npnt, ns = src.shape[1], idx.shape[1], idx.shape[2]
        f = torch.randn(B, C + 3, nin, device=dev)
        alg = B * (npnt * ns * 4 + npnt * ns * (C + 3) * 4 + min(nin, npnt * ns) * (C + 3) * 4)
        ms, best = timeit(lambda: S.grouping_operation(f, idx), flush=flush)
        row(f"grouping C={C + 3} N={nin} P={npnt} S={ns}", ms, best, alg_bytes=alg)
    # larger batch for a meaningful HBM fraction
    for Bb in (32, 64):
        src, q, idx = idxs[1]
        idxb = idx.repeat(Bb // B, 1, 1).contiguous()
        f = torch.randn(Bb, 131, 2048, device=dev)
        alg = Bb * (1024 * 32 * 4 + 1024 * 32 * 131 * 4 + 2048 * 131 * 4)
        ms, best = timeit(lambda: S.grouping_operation(f, idxb), flush=flush)
        row(f"grouping SA2 shape B={Bb}", ms, best, alg_bytes=alg)
        go = torch.randn(Bb, 131, 1024, 32, device=dev)
        f.requires_grad_(True)
        out = S.grouping_operation(f, idxb)
        ms, best = timeit(lambda: torch.autograd.grad(out, f, go, retain_graph=True), flush=flush)
        row(f"grouping bwd SA2 shape B={Bb}", ms, best, alg_bytes=alg)
        del go, out, f

    # ---- three_nn + interpolate (FP1, FP2 shapes, and a large batch)
    for (n, m, C, Bb) in ((512, 256, 256, B), (1024, 512, 256, B), (1024, 512, 256, 64), (1024, 512, 256, 256)):
        u = torch.rand(Bb, n, 3, device=dev) * 6
        k = torch.rand(Bb, m, 3, device=dev) * 6
        ms, best = timeit(lambda: S.three_nn(u, k))
        row(f"three_nn n={n} m={m} B={Bb}", ms, best, alg_bytes=Bb * (n * 12 + m * 12 + n * 24))
        d, i = S.three_nn(u, k)
        w = 1.0 / (d + 1e-8)
        w = (w / w.sum(-1, keepdim=True)).contiguous()
        f = torch.randn(Bb, C, m, device=dev)
        alg = Bb * (n * 3 * 8 + n * C * 4 + m * C * 4)
        ms, best = timeit(lambda: S.three_interpolate(f, i, w), flush=flush)
        row(f"three_interpolate n={n} m={m} C={C} B={Bb}", ms, best, alg_bytes=alg)

    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
