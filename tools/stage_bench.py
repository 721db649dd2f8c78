"""Fused-MLP stages alone at the benchmark's shapes (B = 8 scenes): general kernel vs the shape-specialised one.

    python tools/stage_bench.py [--tpc 1]      -> one line per stage: us, TFLOP/s, fraction of the measured bf16 peak
"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200  # noqa
from sad_b200 import mlp as M


EVENTS = [False]


def t_events(fn, it=20):
    """One launch per CUDA-event pair, every launch queued behind a busy stream (what bench.py's per-call profile sees)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(4000000)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return 1e3 * ts[len(ts) // 2], 1e3 * ts[0]


def t(fn, it=20, reps=10):
    """Device time of one call: `reps` calls captured into a CUDA graph (no host overhead between the launches),
    median / best over `it` replays."""
    if EVENTS[0]:
        return t_events(fn, it)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
        g.replay()
        st.synchronize()
        ts = []
        for _ in range(it):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); g.replay(); b.record(st); st.synchronize()
            ts.append(a.elapsed_time(b) / reps)
    ts.sort()
    return 1e3 * ts[len(ts) // 2], 1e3 * ts[0]


def layers(ch):
    g = torch.Generator().manual_seed(0)
    return M.prepare_layers([((torch.randn(co, ci, generator=g) / ci ** 0.5).cuda(), 0.1 * torch.randn(co, generator=g).cuda())
                             for ci, co in zip(ch[:-1], ch[1:])])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tpc", type=int, default=1)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--no-cf", action="store_true")
    ap.add_argument("--real", action="store_true", help="SA1 / SA2 on real scene data (FPS + ball-query indices)")
    ap.add_argument("--events", action="store_true", help="time single launches with one event pair each (not graph replay)")
    args = ap.parse_args()
    EVENTS[0] = args.events
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        peak = 1400.0
    dev, B = "cuda", args.batch
    M.TILES_PER_CTA[0] = args.tpc
    M._WANT_CF[0] = not args.no_cf
    rows = []
    if args.real:
        from sad_b200 import ops
        from sad_b200.scenes import make_scenes
        xyz_np, feat_np = make_scenes(B, 40000, "surface")
        xyz, feat = torch.from_numpy(xyz_np).to(dev), torch.from_numpy(feat_np).to(dev)
        grid = ops.build_scene_grid(xyz)
        inds = ops.furthest_point_sample(xyz, 2048, grid)
        new_xyz = ops.gather_points(xyz, inds)
        idx = ops.ball_query(0.2, 64, xyz, new_xyz, grid)
        m1 = layers([4, 64, 64, 128])
        hits = (idx != idx[:, :, :1]).sum(-1).float().mean().item() + 1
        for mode, dd in ((False, False), (True, False), (True, True)):
            M.FAST_SA[0], M.DEDUP_SA[0] = mode, dd
            med, best = t(lambda: M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.2, m1))
            print({"stage": "sa1-real", "fast": mode, "dedup": dd, "us": round(med, 1), "mean_distinct_hits": round(hits, 1)}, flush=True)
        M.FAST_SA[0] = True
        f1 = M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.2, m1)
        inds2 = ops.furthest_point_sample(new_xyz, 1024)
        x2 = ops.gather_points(new_xyz, inds2)
        idx2 = ops.ball_query(0.4, 32, new_xyz, x2)
        m2 = layers([131, 128, 128, 256])
        hits2 = (idx2 != idx2[:, :, :1]).sum(-1).float().mean().item() + 1
        for mode, dd in ((False, False), ("single", False), ("single", True), ("pair", False)):
            M.FAST_SA[0], M.DEDUP_SA[0] = mode, dd
            med, best = t(lambda: M.sa_group_mlp(new_xyz, x2, f1, idx2, 0.4, m2))
            print({"stage": "sa2-real", "fast": mode, "dedup": dd, "us": round(med, 1), "mean_distinct_hits": round(hits2, 1)}, flush=True)
        M.DEDUP_SA[0] = True
        # same shapes, random neighbours
        ridx = torch.randint(0, 40000, (B, 2048, 64), device=dev, dtype=torch.int32)
        M.FAST_SA[0] = True
        med, best = t(lambda: M.sa_group_mlp(xyz, new_xyz, feat, ridx, 0.2, m1))
        print({"stage": "sa1-real-xyz-random-idx", "us": round(med, 1)}, flush=True)
        sidx, _ = torch.sort(ridx, dim=-1)
        return
    stages = [("sa1", 40000, 2048, 64, 1, [64, 64, 128]), ("sa2", 2048, 1024, 32, 128, [128, 128, 256]),
              ("sa3", 1024, 512, 16, 256, [128, 128, 256]), ("sa4", 512, 256, 16, 256, [128, 128, 256]),
              ("agg", 1024, 256, 16, 256, [128, 128, 128])]
    for (name, N, P, S, C, hid) in stages:
        xyz = torch.rand(B, N, 3, device=dev)
        new_xyz = torch.rand(B, P, 3, device=dev)
        idx = torch.randint(0, N, (B, P, S), device=dev, dtype=torch.int32)
        feat = torch.randn(B, C, N, device=dev)
        if C >= 64:
            feat._sad_cl = M.to_cl_bf16(feat)
        m = layers([C + 3] + hid)
        flops = 2.0 * B * P * S * sum(a * b for a, b in zip([C + 3] + hid[:-1], hid))
        seen = set()
        for mode in (False, "single", "pair"):
            M.FAST_SA[0] = mode
            inst = M._fast_instance(m, M.sa_layout(C, True), S, P) if mode else -1
            if mode and (inst < 0 or inst in seen):
                continue
            seen.add(inst)
            med, best = t(lambda: M.sa_group_mlp(xyz, new_xyz, feat, idx, 0.3, m))
            rows.append({"stage": name, "kernel": "general" if not mode else f"fast(inst {inst})",
                         "us": round(med, 1), "best_us": round(best, 1), "GFLOP": round(flops / 1e9, 2),
                         "TFLOPs": round(flops / med / 1e6, 1), "frac_of_peak": round(flops / med / 1e6 / peak, 3)})
            print(rows[-1], flush=True)
    M.FAST_SA[0] = True
    # the S == 1 stages (general kernel only, for the total)
    for (name, n, ch, last_relu) in [("fp1", 512, [512, 256, 256], True), ("fp2", 1024, [512, 256, 256], True),
                                      ("vote", 1024, [256, 256, 256, 259], False)]:
        x = torch.randn(B, ch[0], n, device=dev)
        x._sad_cl = M.to_cl_bf16(x)
        m = layers(ch)
        flops = 2.0 * B * n * sum(a * b for a, b in zip(ch[:-1], ch[1:]))
        med, best = t(lambda: M.pointwise_mlp(x, m, last_relu=last_relu))
        rows.append({"stage": name, "kernel": "general", "us": round(med, 1), "best_us": round(best, 1), "GFLOP": round(flops / 1e9, 2),
                     "TFLOPs": round(flops / med / 1e6, 1), "frac_of_peak": round(flops / med / 1e6 / peak, 3)})
        print(rows[-1], flush=True)
    # the specialised point-wise kernel (FP: interpolation fused; voting: residual fused)
    for (name, n, m_) in (("fp1", 512, 256), ("fp2", 1024, 512)):
        kf = torch.randn(B, 256, m_, device=dev); kf._sad_cl = M.to_cl_bf16(kf)
        uf = torch.randn(B, 256, n, device=dev); uf._sad_cl = M.to_cl_bf16(uf)
        idx = torch.randint(0, m_, (B, n, 3), device=dev, dtype=torch.int32)
        w = torch.rand(B, n, 3, device=dev); w = (w / w.sum(-1, keepdim=True)).contiguous()
        m = layers([512, 256, 256])
        flops = 2.0 * B * n * (512 * 256 + 256 * 256)
        for fast in (False, True):
            M.FAST_PW[0] = fast
            med, best = t(lambda: M.fp_interp_mlp(kf, uf, idx, w, m))
            rows.append({"stage": name + "+interp", "kernel": "fast-pw" if fast else "general+interp_cl", "us": round(med, 1),
                         "best_us": round(best, 1), "GFLOP": round(flops / 1e9, 2), "TFLOPs": round(flops / med / 1e6, 1),
                         "frac_of_peak": round(flops / med / 1e6 / peak, 3)})
            print(rows[-1], flush=True)
    sx = torch.rand(B, 1024, 3, device=dev)
    sf = torch.randn(B, 256, 1024, device=dev); sf._sad_cl = M.to_cl_bf16(sf)
    m = layers([256, 256, 256, 259])
    flops = 2.0 * B * 1024 * (256 * 256 * 2 + 256 * 259)
    med, best = t(lambda: M.vote_mlp_fast(sx, sf, m))
    rows.append({"stage": "vote+residual", "kernel": "fast-pw", "us": round(med, 1), "best_us": round(best, 1),
                 "GFLOP": round(flops / 1e9, 2), "TFLOPs": round(flops / med / 1e6, 1), "frac_of_peak": round(flops / med / 1e6 / peak, 3)})
    print(rows[-1], flush=True)
    M.FAST_PW[0] = True
    print(json.dumps({"tpc": args.tpc, "batch": B, "peak_tflops": peak, "rows": rows}))


if __name__ == "__main__":
    main()
