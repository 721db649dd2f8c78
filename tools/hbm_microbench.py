"""HBM-bound drop-in kernels (grouping_operation, three_interpolate) alone, against the measured HBM peak.
Working sets are chosen > L2 (126 MB) where the batch allows; `reps` launches per event pair remove the
Python launch gap from the microsecond-scale kernels.
    python tools/hbm_microbench.py [--out profiles/r01_hbm_microbench.json]"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200 as S  # noqa


def peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def timeit(fn, reps=5, iters=12):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "hbm_microbench.json"))
    a = ap.parse_args()
    hbm, how = peak()
    dev = "cuda:0"
    rows = []
    g = torch.Generator(device="cpu").manual_seed(0)

    def rec(name, ms, alg):
        r = {"op": name, "ms": round(ms, 4), "alg_MB": round(alg / 1e6, 2), "GBps": round(alg / ms / 1e6, 1),
             "hbm_frac": round(alg / ms / 1e6 / hbm, 4), "peak": how}
        rows.append(r)
        print(json.dumps(r), flush=True)

    # grouping_operation at the SA2 / SA3 / SA4 / vote-aggregation shapes (channel-first drop-in surface)
    for (C, N, P, Sn) in ((131, 2048, 1024, 32), (259, 1024, 512, 16), (259, 512, 256, 16), (259, 1024, 256, 16),
                          (4, 40000, 2048, 64)):
        for B in (8, 64):
            if B * C * P * Sn * 4 > 6e9:
                continue
            f = torch.randn(B, C, N, device=dev)
            idx = torch.randint(0, N, (B, P, Sn), generator=g, dtype=torch.int32).to(dev)
            alg = B * (P * Sn * 4 + P * Sn * C * 4 + min(N, P * Sn) * C * 4)
            rec(f"grouping fwd C={C} N={N} P={P} S={Sn} B={B}", timeit(lambda: S.grouping_operation(f, idx)), alg)
            if B == 64 and C == 131:
                go = torch.randn(B, C, P, Sn, device=dev)
                f.requires_grad_(True)
                out = S.grouping_operation(f, idx)
                rec(f"grouping bwd C={C} N={N} P={P} S={Sn} B={B}",
                    timeit(lambda: torch.autograd.grad(out, f, go, retain_graph=True)), alg)
                del go, out
            del f, idx
    # three_interpolate at the FP1 / FP2 shapes
    for (n, m, C) in ((512, 256, 256), (1024, 512, 256)):
        for B in (8, 64, 256):
            u = torch.rand(B, n, 3, device=dev) * 6
            k = torch.rand(B, m, 3, device=dev) * 6
            d, i = S.three_nn(u, k)
            w = 1.0 / (d + 1e-8)
            w = (w / w.sum(-1, keepdim=True)).contiguous()
            f = torch.randn(B, C, m, device=dev)
            alg = B * (n * 3 * 8 + n * C * 4 + m * C * 4)
            rec(f"three_interpolate fwd n={n} m={m} C={C} B={B}", timeit(lambda: S.three_interpolate(f, i, w)), alg)
    # channel-last bf16 interpolation (what the fused product path uses)
    import ctypes
    from sad_b200 import _lib
    lib = _lib.load()
    for (n, m, C) in ((512, 256, 256), (1024, 512, 256)):
        for B in (8, 64, 512):
            u = torch.rand(B, n, 3, device=dev) * 6
            k = torch.rand(B, m, 3, device=dev) * 6
            d, i = S.three_nn(u, k)
            w = 1.0 / (d + 1e-8)
            w = (w / w.sum(-1, keepdim=True)).contiguous()
            fcl = torch.randn(B, m, C, device=dev).to(torch.bfloat16)
            o = torch.empty(B, n, C, device=dev, dtype=torch.bfloat16)
            vp = ctypes.c_void_p
            fn = lambda: lib.sad_three_interpolate_cl_fwd(B, C, m, n, vp(fcl.data_ptr()), vp(i.data_ptr()), vp(w.data_ptr()),
                                                          vp(o.data_ptr()), vp(torch.cuda.current_stream().cuda_stream))
            alg = B * (n * 3 * 8 + n * C * 2 + m * C * 2)
            rec(f"three_interpolate_cl (bf16, channel-last) n={n} m={m} C={C} B={B}", timeit(fn, reps=10), alg)
    json.dump({"hbm_peak_gbs": hbm, "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
