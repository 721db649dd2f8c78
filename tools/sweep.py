"""BASELINE config 5: scenes/s of the hot path over points-per-scene x batch (one GPU; the path shards by scene,
so N GPUs multiply it -- measured for the headline config in bench.py), next to the CPU path on a subset.
    python tools/sweep.py --out profiles/r01_sweep.json"""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200  # noqa
from sad_b200.config import LAYER_CFG, make_params
from sad_b200.engine import PipelinedHotPath
from sad_b200.modules import SADHotPath
from sad_b200.scenes import make_scenes, make_sizes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--slots", type=int, default=16)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    model = SADHotPath(1).load_params(make_params(0)).to(dev).eval()
    rows = []
    for N in (20000, 40000, 80000, 120000, 200000):
        for B in (1, 8, 32, 64):
            if B * N > 3_300_000:          # keep the sweep within a few GB / seconds
                continue
            xyz, feat = make_scenes(B, N, "surface")
            size = make_sizes(B, LAYER_CFG["agg"][0])
            d = tuple(torch.from_numpy(t).to(dev) for t in (xyz, feat, size))
            eng = PipelinedHotPath(model, B, N, slots=a.slots, device=dev)
            for _ in range(a.slots):
                eng.submit_device(*d)
            eng.drain()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            main_s = torch.cuda.current_stream(dev)
            e0.record(main_s)
            for k in range(a.steps):
                eng.submit_device(*d, after=e0 if k < eng.slots else None)
            eng.join(main_s)
            e1.record(main_s)
            torch.cuda.synchronize()
            eng.drain()
            ms = e0.elapsed_time(e1) / a.steps
            # latency of one batch alone
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(main_s)
            eng.submit_device(*d, after=s0)
            eng.join(main_s)
            s1.record(main_s)
            torch.cuda.synchronize()
            eng.drain()
            r = {"points_per_scene": N, "batch": B, "scenes_per_s": round(B / ms * 1e3, 1), "ms_per_batch_pipelined": round(ms, 4),
                 "batch_latency_ms": round(s0.elapsed_time(s1), 4)}
            if (N, B) in ((20000, 1), (40000, 8)):
                from bench import cpu_hot_path_rate, host_threads
                cpu_hot_path_rate(1, 4000)
                rate, secs, thr = cpu_hot_path_rate(B, N, reps=1)
                r["cpu_scenes_per_s"] = round(rate, 3)
                r["cpu_threads"] = thr
            rows.append(r)
            print(json.dumps(r), flush=True)
            del eng
            torch.cuda.empty_cache()
    json.dump({"note": f"one B200, {a.slots} batches in flight (throughput FPS policy), inputs resident; surface scenes", "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
