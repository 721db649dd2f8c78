#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the 3DSAD hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one pass of the hot path (4 SA + 2 FP backbone -> voting -> size-adaptive vote
aggregation) over one batch of synthetic scenes.  Workload at N=1 = BASELINE.json
configs[1]: B=8 scenes x 40k points per GPU (weak scaling: every rank gets its own 8
scenes, no data-path collective).  metric = scenes/s, whole job.

  value : inputs resident in HBM; K batches through the pipelined executor (engine.py: one CUDA
          graph + stream per slot, `--slots` batches in flight), one CUDA-event pair around all K;
          32 rotating input sets (> L2) instead of an L2 flush.  config.batch_latency_ms is one
          batch alone on the GPU.
  e2e   : the same metric through PipelinedHotPath.submit_host / result with HOST (pinned)
          buffers: H2D of xyz/features/sizes and D2H of the cluster centres + features inside
          the timed region, host wall clock, every result read on the host.
  roofline     : the fused gather + MLP + max-pool launches (tensor roofline; they own the largest share of the step's
                 SM-time), measured live with CUDA events; `fps` = the sampling chain in picks/s; `hbm_kernels` = the
                 grouping / interpolation kernels against the HBM roofline at B = 64 / 256.
  cpu_baseline : the oracle's C/OpenMP port + NumPy MLP ("port") on the box's host cores,
                 bounded sample, rank 0 at N=1 only.
  --impl reference : the reference arm.  The mounted reference is a README (no code), so
          per BASELINE.json the CPU oracle port IS the reference implementation of the path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "scenes/sec (40k pts) backbone+size-adaptive clustering"
UNIT = "scenes/s"
B_PER_GPU = 8
N_POINTS = 40000
WORKLOAD = "configs[1]: VoteNet-style backbone (4 SA + 2 FP) + voting + size-adaptive vote aggregation, " \
           "B=8 x 40k-point synthetic ScanNet-shape (surface) scenes per GPU"


def workload_config(n_gpus):
    """`config` of the JSON line: the workload only, identical in both arms (what differs between the arms -- pipeline
    depth, FPS scheduling, dtypes of the MLP -- is in `run`)."""
    wl = WORKLOAD if (B_PER_GPU, N_POINTS) == (8, 40000) else \
        f"configs[4] sweep point: the configs[1] path at B={B_PER_GPU} x {N_POINTS}-point scenes per GPU (--batch / --points)"
    return {"workload": wl, "scenes_per_gpu_per_step": B_PER_GPU, "points_per_scene": N_POINTS,
            "scene_kind": "surface (floor + walls + 12 boxes, 5 mm noise), seeded", "input_feature": "height (1 channel)",
            "layers": "SA1 2048x64 r0.2 [4,64,64,128] | SA2 1024x32 r0.4 [131,128,128,256] | SA3 512x16 r0.8 [259,128,128,256] | "
                      "SA4 256x16 r1.2 [259,128,128,256] | FP1,FP2 [512,256,256] | vote [256,256,256,259] | "
                      "agg 256x16 adaptive radius [259,128,128,128]",
            "parallelism": f"scene-data-parallel x{n_gpus}, no collective"}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def measured_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm": d["hbm_gbs"], "tensor": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tensor": 1400.0, "src": "fallback"}


# ----------------------------------------------------------------------------- CPU arm
def cpu_hot_path_rate(n_scenes, n_points, reps=1, seed0=0):
    """Oracle C/OpenMP port + NumPy MLP over `n_scenes` scenes; returns (scenes/s, seconds, threads)."""
    import numpy as np  # noqa: F401
    from oracle import sad_oracle as O, c_port as C
    import sad_b200  # noqa: F401  (package import only: config + scene generator, no kernels)
    from sad_b200.config import LAYER_CFG, make_params
    from sad_b200.scenes import make_scenes, make_sizes

    C.build()
    C.set_threads(host_threads())
    params = make_params(0)
    xyz, feat = make_scenes(n_scenes, n_points, "surface", first_scene=seed0)
    size = make_sizes(n_scenes, LAYER_CFG["agg"][0], first_scene=seed0)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        O.detector_hot_path(xyz, feat, size, params, LAYER_CFG, impl=C)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_scenes / best, best, host_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    n_scenes = max(1, min(B_PER_GPU, threads))          # FPS parallelises over scenes only
    cpu_hot_path_rate(1, 4000)                           # warm caches / OpenMP pool
    # bounded sample: keep the whole run within ~150 s of CPU time
    budget = 150.0 / max(1, args.warmup + args.steps)
    _, dt, _ = cpu_hot_path_rate(n_scenes, N_POINTS, reps=1, seed0=7)
    if dt > budget:
        n_scenes = max(1, int(n_scenes * budget / dt))
    times = []
    for s in range(args.warmup + args.steps):
        rate, dt, _ = cpu_hot_path_rate(n_scenes, N_POINTS, reps=1, seed0=100 * s)
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n_scenes * len(times) / total
    sample = f"{n_scenes} scenes x {N_POINTS} pts per step (one OpenMP pass over all {threads} host threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / len(times), 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "run": {"note": "reference mount is README-only; the CPU oracle port (C/OpenMP + NumPy MLP, fp32) is the reference path",
                "fps_threads": f"FPS parallelises over scenes only: {n_scenes} of the {threads} host threads busy in the dominant phase"},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock + throttle reasons through NVML every ~2 ms on a side thread
    (the recipe's nvidia-smi line, without its start-up latency)."""

    def __init__(self, index):
        self.rows, self.ok, self._stop = [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def _poll(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.rows.append((time.perf_counter(), sm, rs))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def stop(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self._stop = True
        self.th.join(timeout=1.0)
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        sm, reasons = [], set()
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        if not rows and self.rows:      # a window shorter than one polling period: the sample nearest to it
            mid = 0.5 * (t0 + t1)
            rows = [min(self.rows, key=lambda r: abs(r[0] - mid))]
        for (t, c, rs) in rows:
            sm.append(c)
            for name, bit in bits.items():
                if rs & bit:
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- roofline
def call_cost(name, a):
    """(algorithmic bytes, flops) of one C-ABI call from its integer arguments (SURVEY 8(d) formulae)."""
    v = [x if isinstance(x, int) else None for x in a]
    if name in ("sad_furthest_point_sample_fwd", "sad_furthest_point_sample_grid_fwd"):
        B, N, P = v[0], v[1], v[2]
        return B * (N * 12 + P * 4), 0
    if name == "sad_scene_grid_build":
        B, N = v[0], v[1]
        return B * (N * 12 + N * 16 + 32769 * 4), 0
    if name == "sad_ball_query_grid_fwd":
        B, N, P, S = v[0], v[1], v[2], v[5]
        return B * (N * 12 + P * 12 + P * S * 4), 0
    if name in ("sad_ball_query_fwd", "sad_ball_query_adaptive_fwd"):
        B, N, P, S = v[0], v[1], v[2], v[4]
        return B * (N * 12 + P * 12 + P * S * 4), 0
    if name == "sad_grouping_operation_fwd":
        B, C, N, P, S = v[:5]
        return B * (P * S * 4 + P * S * C * 4 + min(N, P * S) * C * 4), 0
    if name == "sad_gather_operation_fwd":
        B, C, N, P = v[:4]
        return B * (P * 4 + P * C * 4 + min(N, P) * C * 4), 0
    if name == "sad_three_nn_fwd":
        B, n, m = v[:3]
        return B * (n * 12 + m * 12 + n * 24), 0
    if name == "sad_three_interpolate_fwd":
        B, C, m, n = v[:4]
        return B * (n * 24 + n * C * 4 + m * C * 4), 0
    if name == "sad_three_interpolate_cl_fwd":
        B, C, m, n = v[:4]
        return B * (n * 24 + n * C * 2 + m * C * 2), 0
    if name == "sad_cf_to_cl_bf16":
        B, C, N = v[:3]
        return B * C * N * 6, 0
    if name in ("sad_sa_mlp_fwd", "sad_sa_mlp_dedup_fwd"):
        # specialised fused SA stage: (inst, B, N, P, feat_cl, xyz, xyzw, new_xyz, idx, radius, radius_t, norm, extra, E, ...)
        from sad_b200 import _lib as _L
        import ctypes as _ct
        info = (_ct.c_int * 5)()
        _L.load().sad_sa_mlp_instance_info(int(a[0]), info)
        _, NF, H, _, S = list(info)
        B, N, P, E, c3 = v[1], v[2], v[3], v[13], v[16]
        rows = B * P * S
        cin0 = NF * 64 + 3 + E
        flops = 2 * rows * (cin0 * H + H * H + H * c3)
        nbytes = rows * 4 + B * min(N, P * S) * (NF * 128 + 12 + E * 4) + B * P * c3 * 6
        return nbytes, flops
    if name == "sad_pw_mlp_fwd":
        # specialised fused point-wise stage: (kind, B, n, m, src_cl, known_cl, nn_idx, nn_w, ..., c_last [12], out_cf, out_cl, ...)
        kind, B, n, m, c_last = v[0], v[1], v[2], v[3], v[12]
        rows = B * n

        def live(x):
            return getattr(x, "value", x) not in (None, 0)
        outs = (4 if live(a[13]) else 0) + (2 if live(a[14]) else 0)
        if kind == 0:          # FP: [interp(known) 256 | skip 256] -> 256 -> 256
            flops = 2 * rows * (512 * 256 + 256 * c_last)
            nbytes = rows * (24 + 256 * 2 + c_last * outs) + B * m * 256 * 2
        else:                  # voting: 256 -> 256 -> 256 -> 3 + 256, vote = seed + y
            flops = 2 * rows * (256 * 256 * 2 + 256 * c_last)
            nbytes = rows * (256 * 2 + 12 + 256 * 4 + 12 + 256 * outs)
        return nbytes, flops
    if name == "sad_mlp_tf32_fwd":
        # tf32 fused stage: (B, N, P, S, known, m, CI, nn_idx, nn_w, feat, CF, idx, xyz, ..., E [18], n_layers, w, b, c_out [22], ...)
        B, N, P, S, CI, CF, E = v[0], v[1], v[2], v[3], v[6], v[10], v[18]
        has_xyz = getattr(a[12], "value", a[12]) not in (None, 0)
        cout = list(a[22])
        rows = B * P * S
        cin0 = CI + CF + (3 + E if has_xyz else 0)
        flops = 2 * rows * sum(ci * co for ci, co in zip([cin0] + cout[:-1], cout))
        nbytes = rows * 4 + B * min(N, P * S) * (CF * 4 + (12 + E * 4 if has_xyz else 0)) + B * P * cout[-1] * 8
        return nbytes, flops
    if name == "sad_shared_mlp_fwd":
        # fused stage: bytes = idx + distinct gathered rows (bf16) + outputs; flops = 2*rows*sum(Cin*Cout)
        B, N, P, S, C0, C1in, E, nl = v[0], v[1], v[2], v[3], v[5], v[7], v[15], v[16]
        cout = list(a[19])
        has_xyz = bool(a[8].value) if hasattr(a[8], "value") else bool(a[8])
        cin0 = C0 + C1in + (3 if has_xyz else 0) + E
        rows = B * P * S
        flops = 2 * rows * sum(ci * co for ci, co in zip([cin0] + cout[:-1], cout))
        nbytes = rows * 4 + B * min(N, P * S) * (C0 * 2 + 12 + E * 4) + B * P * C1in * 2 + B * P * cout[-1] * 6
        return nbytes, flops
    return 0, 0


MLP_DTYPE = ["bf16"]
MLP_REPEAT = 8      # fused-MLP launches are re-issued this many times inside one CUDA-event pair (same inputs, same outputs)


def build_roofline(model, xyz, feat, size, reps=3):
    import torch
    from sad_b200 import _lib
    peaks = measured_peaks()
    if MLP_DTYPE[0] == "tf32":      # kind::tf32 runs at half the kind::f16 rate; only the bf16 peak is measured on this pool
        peaks = dict(peaks, tensor=peaks["tensor"] / 2, src=peaks["src"] + " (bf16 sustained / 2 for tf32)")
    agg = {}
    stages = {}
    for _ in range(reps):
        with _lib.CallProfiler(repeat={"sad_sa_mlp_fwd": MLP_REPEAT, "sad_sa_mlp_dedup_fwd": MLP_REPEAT, "sad_pw_mlp_fwd": MLP_REPEAT,
                                       "sad_shared_mlp_fwd": MLP_REPEAT, "sad_mlp_tf32_fwd": MLP_REPEAT}) as prof:
            with torch.no_grad():
                torch.cuda._sleep(60000000)      # ~30 ms: every launch of the forward is queued before the first one runs
                model(xyz, feat, size)
        for name, a, ms in prof.rows():
            name = name.replace("_grid_policy_fwd", "_grid_fwd")      # same op, explicit scheduling policy
            name = name.replace("_prefix_fwd", "_fwd")                # same op over prefix-ordered input
            nbytes, flops = call_cost(name, a)
            key = name.replace("sad_", "").replace("_fwd", "")
            picks = 0
            if name in ("sad_furthest_point_sample_fwd", "sad_furthest_point_sample_grid_fwd"):
                key += f"[N={a[1]}]"
                picks = int(a[2])
            d = agg.setdefault(key, {"ms": 0.0, "bytes": 0, "flops": 0, "launches": 0, "picks": picks})
            if flops:      # one row per fused-MLP launch, in call order (SA1..SA4, FP1, FP2, voting, aggregation)
                sk = (name, tuple(x for x in a[:11] if isinstance(x, int)))
                st = stages.setdefault(sk, {"ms": 0.0, "flops": flops, "order": len(stages)})
                st["ms"] += ms / reps
            d["ms"] += ms / reps
            d["bytes"] += nbytes / reps
            d["flops"] += flops / reps
            d["launches"] += 1.0 / reps
    kernels = []
    for key, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        gbs = d["bytes"] / d["ms"] / 1e6 if d["ms"] > 0 else 0.0
        row = {"kernel": key, "ms_per_step": round(d["ms"], 4), "launches_per_step": round(d["launches"], 1),
               "alg_MB_per_step": round(d["bytes"] / 1e6, 3), "GBps": round(gbs, 1),
               "hbm_frac": round(gbs / peaks["hbm"], 4)}
        if d["flops"]:
            tfs = d["flops"] / d["ms"] / 1e9 if d["ms"] > 0 else 0.0
            row.update(GFLOP_per_step=round(d["flops"] / 1e9, 2), TFLOPs=round(tfs, 1),
                       tensor_frac=round(tfs / peaks["tensor"], 4))
        kernels.append(row)
    # ---- headline: the fused gather + MLP + max-pool launches (tensor roofline).  They own the largest share of the
    # step's SM-time; the FPS chain is a serial-latency kernel and is reported in its honest unit below.
    mlp_rows = [k for k in kernels if k["kernel"] in ("shared_mlp", "sa_mlp", "sa_mlp_dedup", "pw_mlp", "mlp_tf32")]
    flops = sum(agg[k["kernel"]]["flops"] for k in mlp_rows)
    ms = sum(agg[k["kernel"]]["ms"] for k in mlp_rows)
    n_l = sum(agg[k["kernel"]]["launches"] for k in mlp_rows)
    tfs = flops / ms / 1e9 if ms > 0 else 0.0
    roof = {"bound": "tensor", "kernel": f"fused gather + shared MLP + max-pool ({n_l:.0f} launches per step: "
                                         + ("SA1-SA4 + aggregation [sa_mlp_kernel], FP1/FP2/voting [pw_mlp_kernel])" if MLP_DTYPE[0] == "bf16"
                                            else "mlp_tf32_kernel, fp32 activations, kind::tf32)"),
            "achieved": round(tfs, 1), "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": round(tfs / peaks["tensor"], 4),
            "traffic": None, "peak_source": peaks["src"], "alg_flops_per_launch": round(flops / max(1.0, n_l)),
            "launch_ms": round(ms / max(1.0, n_l), 4), "ms_per_step_all_launches": round(ms, 4),
            "note": ("bf16 operands, fp32 accumulate; peak = sustained cuBLAS bf16 (MEASURED_PEAKS.json)" if MLP_DTYPE[0] == "bf16"
                     else "tf32 operands, fp32 accumulate; peak = half the sustained cuBLAS bf16 figure (MEASURED_PEAKS.json)")
                    + "; algorithmic flops = "
                    "2 * rows * sum(Cin*Cout) with the real (unpadded) channel counts; device time by CUDA events, launches "
                    "queued behind a busy stream so no host gap is inside an event pair; each fused-MLP call is issued "
                    f"{MLP_REPEAT} times back to back inside its event pair (idempotent: same inputs and outputs) and the "
                    "elapsed time divided, because one event pair around a single 148-CTA launch adds 6-19 us "
                    "(sa_mlp_dedup = SA1 / SA2 run duplicate-free: the flops are the algorithmic ones of SURVEY 8(d), every "
                    "(point, sample) row counted; the launch, plan kernel included, executes only the rows that are not "
                    "padding copies of a neighbourhood's first hit) "
                    "(tools/stage_bench.py --events vs graph replay, profiles/r02_stage_bench_*.txt)"}
    per_stage = []
    for k in mlp_rows:
        per_stage.append({"kernel": k["kernel"], "ms": k["ms_per_step"], "GFLOP": k.get("GFLOP_per_step"),
                          "TFLOPs": k.get("TFLOPs"), "frac": k.get("tensor_frac")})
    roof["by_kernel"] = per_stage
    names = ["SA1", "SA2", "SA3", "SA4", "FP1", "FP2", "voting", "aggregation"]
    roof["by_launch"] = []
    for i, (sk, st) in enumerate(sorted(stages.items(), key=lambda kv: kv[1]["order"])):
        t = st["flops"] / st["ms"] / 1e9 if st["ms"] > 0 else 0.0
        roof["by_launch"].append({"stage": names[i] if i < len(names) else str(i), "entry": sk[0].replace("sad_", ""),
                                  "dims": list(sk[1][:4]), "us": round(1e3 * st["ms"], 1), "GFLOP": round(st["flops"] / 1e9, 2),
                                  "TFLOPs": round(t, 1), "frac": round(t / peaks["tensor"], 4)})
    # ---- FPS in its own unit: dependent picks per second
    fps_rows = []
    for k in kernels:
        if k["kernel"].startswith("furthest_point_sample"):
            d = agg[k["kernel"]]
            picks = d.get("picks", 0)
            fps_rows.append({"kernel": k["kernel"], "ms": k["ms_per_step"], "picks_per_scene": picks,
                             "us_per_pick": round(1e3 * d["ms"] / max(1, picks), 4),
                             "picks_per_s_per_scene": round(picks / (d["ms"] / 1e3)) if d["ms"] > 0 else None})
    return roof, kernels, fps_rows


def hbm_kernel_block(dev):
    """north_star's HBM bar, measured in this run: grouping_operation fwd at the SA2/SA3/SA4/vote shapes and
    three_interpolate fwd (drop-in fp32 and the channel-last bf16 variant of the product path) at the FP1/FP2
    shapes, B = 64 and 256 scenes per launch (working sets > L2), `reps` launches per event pair."""
    import ctypes
    import torch
    import sad_b200 as S
    from sad_b200 import _lib
    peaks = measured_peaks()
    g = torch.Generator(device="cpu").manual_seed(0)
    rows = []

    def timeit(fn, reps=4, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(2000000)           # launches queue behind a busy stream: no host gap inside the pair
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / reps)
        ts.sort()
        return ts[len(ts) // 2]

    def rec(name, ms, alg):
        gbs = alg / ms / 1e6
        rows.append({"op": name, "ms": round(ms, 4), "alg_MB": round(alg / 1e6, 1), "GBps": round(gbs, 1),
                     "frac": round(gbs / peaks["hbm"], 4)})

    for (tag, C, N, P, Sn) in (("SA2", 131, 2048, 1024, 32), ("SA3", 259, 1024, 512, 16), ("SA4", 259, 512, 256, 16),
                               ("vote", 259, 1024, 256, 16)):
        for B in (64, 256):
            if B * C * P * Sn * 4 > 5e9:
                continue
            f = torch.randn(B, C, N, device=dev)
            idx = torch.randint(0, N, (B, P, Sn), generator=g, dtype=torch.int32).to(dev)
            alg = B * (P * Sn * 4 + P * Sn * C * 4 + min(N, P * Sn) * C * 4)
            rec(f"grouping_operation fwd {tag} shape B={B}", timeit(lambda: S.grouping_operation(f, idx)), alg)
            del f, idx
    lib = _lib.load()
    vp = ctypes.c_void_p
    for (tag, n, m, C) in (("FP1", 512, 256, 256), ("FP2", 1024, 512, 256)):
        for B in (64, 256):
            u = torch.rand(B, n, 3, device=dev) * 6
            k = torch.rand(B, m, 3, device=dev) * 6
            d, i = S.three_nn(u, k)
            w = 1.0 / (d + 1e-8)
            w = (w / w.sum(-1, keepdim=True)).contiguous()
            f = torch.randn(B, C, m, device=dev)
            rec(f"three_interpolate fwd {tag} shape B={B}", timeit(lambda: S.three_interpolate(f, i, w)),
                B * (n * 3 * 8 + n * C * 4 + m * C * 4))
            fcl = f.transpose(1, 2).contiguous().to(torch.bfloat16)
            o = torch.empty(B, n, C, device=dev, dtype=torch.bfloat16)
            st = vp(torch.cuda.current_stream().cuda_stream)
            rec(f"three_interpolate_cl (bf16 channel-last) fwd {tag} shape B={B}",
                timeit(lambda: lib.sad_three_interpolate_cl_fwd(B, C, m, n, vp(fcl.data_ptr()), vp(i.data_ptr()),
                                                                vp(w.data_ptr()), vp(o.data_ptr()), st)),
                B * (n * 3 * 8 + n * C * 2 + m * C * 2))
            del u, k, d, i, w, f, fcl, o
    return {"peak_GBps": peaks["hbm"], "peak_source": peaks["src"], "bar": "north_star: >= 0.60 of the HBM roofline on the "
            "grouping / interpolation kernels", "rows": rows}


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    import sad_b200 as S  # noqa: F401
    from sad_b200 import _lib
    from sad_b200.config import LAYER_CFG, make_params
    from sad_b200.engine import PipelinedHotPath
    from sad_b200.modules import SADHotPath
    from sad_b200.scenes import make_scenes, make_sizes

    _lib.load()
    MLP_DTYPE[0] = args.mlp_dtype
    model = SADHotPath(input_feature_dim=1, mlp_dtype=args.mlp_dtype).load_params(make_params(0)).to(dev).eval()

    # distinct scenes per rank and per rotating input set; the sets together exceed L2 (126 MB),
    # so no step finds its inputs cached from an earlier one
    from sad_b200 import dist as D          # the package's scene-sharding / max-over-ranks bookkeeping
    NSETS = args.sets
    sets = []
    for s in range(NSETS):
        first = D.weak_scaling_first_scene(rank, B_PER_GPU, input_sets=NSETS, set_index=s)
        xyz, feat = make_scenes(B_PER_GPU, N_POINTS, "surface", first_scene=first)
        size = make_sizes(B_PER_GPU, LAYER_CFG["agg"][0], first_scene=first)
        host = tuple(torch.from_numpy(a).pin_memory() for a in (xyz, feat, size))
        sets.append({"host": host, "dev": tuple(h.to(dev) for h in host)})
    set_bytes = sum(int(h.numel() * h.element_size()) for h in sets[0]["host"])

    eng = PipelinedHotPath(model, B_PER_GPU, N_POINTS, feat_dim=1, slots=args.slots, device=dev,
                           fps_policy=args.fps_policy, mlp_tiles_per_cta=args.tpc, native_submit=not args.py_submit,
                           linear_graph=not args.forked_graph)
    main = torch.cuda.current_stream(dev)

    def barrier():
        D.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM, K batches through the pipelined executor, device-timed.  Each slot's static
    # input buffers hold one of the rotating input sets (written before the timed region), so a step is the launch of a
    # slot's graph and nothing else; when the slots together would fit L2 the sets are copied in per step instead.
    resident = eng.slots * set_bytes > 1.2 * 126e6 and not args.copy_inputs
    if resident:
        for i in range(eng.slots):
            for dst, src in zip(eng.slot_inputs(i), sets[i % NSETS]["dev"]):
                dst.copy_(src)
        torch.cuda.synchronize()

    def submit_value(k, after=None):
        if resident:
            return eng.submit_resident(after=after)
        return eng.submit_device(*sets[k % NSETS]["dev"], after=after)

    for w in range(args.warmup):
        submit_value(w)
    eng.drain()
    barrier()
    sampler = ClockSampler(local) if rank == 0 and not args.no_clocks else None
    t_wall0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(main)
    t_sub0 = time.perf_counter()
    sub_t = []
    for k in range(args.steps):
        submit_value(args.warmup + k, after=ev0 if k < eng.slots else None)
        sub_t.append(time.perf_counter())
    # host time to queue one batch: the calls that cannot block on back-pressure (the first `slots` of them)
    n_free = max(1, min(args.steps, eng.slots))
    sub_first = sorted((b - a) * 1e6 for a, b in zip([t_sub0] + sub_t[:n_free - 1], sub_t[:n_free]))
    host_submit_us = sub_first[len(sub_first) // 2]
    eng.join(main)
    ev1.record(main)
    barrier()
    eng.drain()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    if clocks is None and rank == 0:
        clocks = {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["not sampled (--no-clocks: tools only, not a bench value)"]}
    total_ms = ev0.elapsed_time(ev1)

    # ---- latency of ONE batch in isolation (nothing else in flight), device-timed
    lat = []
    for k in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        eng.submit_device(*sets[k % NSETS]["dev"], after=a)
        eng.join(main)
        b.record(main)
        torch.cuda.synchronize()
        eng.drain()
        lat.append(a.elapsed_time(b))
    lat.sort()

    # ---- e2e: host (pinned) buffers in, host results out, every step; host wall clock
    checksum = 0.0
    for w in range(max(1, args.warmup // 2)):
        eng.result(eng.submit_host(*sets[w % NSETS]["host"]))
    barrier()
    t0 = time.perf_counter()
    tickets = []
    for k in range(args.steps):
        tickets.append(eng.submit_host(*sets[(args.warmup + k) % NSETS]["host"]))
        if len(tickets) >= eng.slots:                     # results are consumed in order, one pipeline depth behind
            cx, cf = eng.result(tickets.pop(0))
            checksum += float(cf[0, 0, 0]) + float(cx[0, 0, 0])
    for t in tickets:
        cx, cf = eng.result(t)
        checksum += float(cf[0, 0, 0]) + float(cx[0, 0, 0])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    out_host = eng.result(0)
    h2d = set_bytes
    d2h = sum(int(o.numel() * o.element_size()) for o in out_host)

    # ---- max over ranks
    total_ms, e2e_ms = D.reduce_scalars([total_ms, e2e_s * 1e3], "max", device=dev)
    scenes = B_PER_GPU * world * args.steps
    value = scenes / (total_ms / 1e3)
    e2e_value = scenes / (e2e_ms / 1e3)

    if rank == 0:
        model.backbone.overlap_geometry = False          # per-kernel view: one stream, nothing overlapped
        from sad_b200 import modules as _modules
        _modules.FPS_POLICY[0] = args.fps_policy         # the same FPS kernel the captured graphs run
        roof, kernels, fps_rows = build_roofline(model, *sets[0]["dev"])
        _modules.FPS_POLICY[0] = "latency"
        model.backbone.overlap_geometry = True
        roof = attach_traffic(roof)
        hbm_block = hbm_kernel_block(dev) if world == 1 and not args.no_hbm else None
        cpu = None
        if world == 1 and not args.no_cpu:
            n_s = max(1, min(B_PER_GPU, host_threads()))
            cpu_hot_path_rate(1, 4000)
            rate, secs, threads = cpu_hot_path_rate(n_s, N_POINTS, reps=2)
            cpu = {"value": round(rate, 4), "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{n_s} scenes x {N_POINTS} pts, best of 2 passes ({secs:.2f} s per pass), "
                             "oracle C/OpenMP port + NumPy(BLAS) MLP"}
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.mlp_dtype, "data": "synthetic",
            "config": workload_config(world),
            "run": {"l2": (f"{eng.slots} input sets x {set_bytes / 1e6:.1f} MB = {eng.slots * set_bytes / 1e6:.0f} MB > 126 MB L2, "
                           "one resident in each slot's input buffers (inputs larger than L2, no flush; value = graph "
                           "launches only, no input copy)") if resident else
                          (f"{NSETS} rotating input sets x {set_bytes / 1e6:.1f} MB = {NSETS * set_bytes / 1e6:.0f} MB "
                           "> 126 MB L2 (inputs larger than L2, no flush), copied device-to-device into the slot per step"),
                    "pipeline": f"{eng.slots} batches in flight (one CUDA graph + stream per slot); every batch runs "
                                "the full path and results are delivered in order",
                    "steady_state": bool(args.steps >= 2 * eng.slots),
                    "steady_state_note": "with steps < 2 x slots the timed window is one pipeline fill and drain: every "
                                         "batch is submitted at once and `value` is total work / total time, not the "
                                         "sustained rate (python bench.py without flags runs 200 steps)",
                    "fps_policy": f"{args.fps_policy} (throughput = one SM per scene for the 40k-point FPS, latency = "
                                  "4-SM cluster per scene; identical indices)",
                    "submit": ("one sad_engine_submit call per batch (C ABI)" if eng.native_submit else "PyTorch calls"),
                    "graph": ("straight-line (one stream per batch)" if eng.linear_graph else "forked (side streams inside the batch)"),
                    "host_submit_us_per_batch": round(host_submit_us, 1),
                    "batch_latency_loaded_ms": round(eng.slots * (total_ms / args.steps), 4),
                    "batch_latency_ms": round(lat[len(lat) // 2], 4),
                    "search_dtype": "f32 (bit-exact indices)", "mlp_dtype": f"{args.mlp_dtype} in / f32 accumulate"},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_ms / args.steps, 4), "checksum": round(checksum, 4)},
            "gpu_launches": int(eng.launches_per_batch * args.steps),
            "gpu_launches_per_step": eng.launches_per_batch,
            "roofline": roof, "fps": fps_rows, "kernels": kernels, "clocks": clocks,
        }
        if hbm_block is not None:
            line["hbm_kernels"] = hbm_block
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def attach_traffic(roof):
    """DRAM traffic of the dominant kernel from the committed `ncu --set full` summary
    (profiles/ncu_full_summary.json, written by tools/ncu_full_summary.py), per launch."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_full_summary.json")))
        rows = [r for r in d.get("kernels", []) if r.get("roofline_key") in ("shared_mlp", "sa_mlp", "pw_mlp") and "dram_bytes_per_launch" in r]
        if rows:
            roof["traffic"] = int(sum(r["dram_bytes_per_launch"] for r in rows) / len(rows))
            roof["traffic_source"] = "profiles/ncu_full_summary.json (ncu --set full, mean over the step's launches of this kernel)"
    except Exception:  # noqa: BLE001
        pass
    return roof


_RESULT_FD = 1


def emit(line):
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--slots", type=int, default=32, help="batches in flight in the pipelined executor")
    ap.add_argument("--fps-policy", default="throughput", choices=["throughput", "throughput_paired", "latency"],
                    help="scheduling of the 40k-point FPS (same indices either way)")
    ap.add_argument("--sets", type=int, default=32, help="rotating input sets (32 x 5.1 MB > L2)")
    ap.add_argument("--tpc", type=int, default=6, help="fused-MLP tiles per CTA under the pipelined executor (scheduling only)")
    ap.add_argument("--batch", type=int, default=8, help="scenes per GPU per step (default = configs[1]; other values = sweep points)")
    ap.add_argument("--points", type=int, default=40000, help="points per scene (default = configs[1])")
    ap.add_argument("--mlp-dtype", default="bf16", choices=["bf16", "tf32"],
                    help="operand precision of the fused MLP stages (tf32: fp32 activations, csrc/mlp_tf32.cu)")
    ap.add_argument("--forked-graph", action="store_true", help="capture the coordinate-only chain on side streams (forked graph; comparison)")
    ap.add_argument("--py-submit", action="store_true", help="queue batches through PyTorch calls instead of sad_engine_submit (comparison)")
    ap.add_argument("--copy-inputs", action="store_true", help="value: copy every step's inputs device-to-device into the slot (submit_device) "
                    "instead of keeping one input set resident per slot")
    ap.add_argument("--no-clocks", action="store_true", help="tools: no NVML clock sampling thread during the timed region")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-hbm", action="store_true", help="skip the hbm_kernels micro-benchmark block")
    args = ap.parse_args()
    global B_PER_GPU, N_POINTS
    B_PER_GPU, N_POINTS = args.batch, args.points
    # stdout carries exactly ONE JSON line: anything a library prints to fd 1 during the run (NCCL's version banner
    # under NCCL_DEBUG=VERSION, for one) is sent to stderr, and the line is written to the original stdout
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        # torchrun pins its workers to one OpenMP thread; the CPU arm uses every host thread it can get
        for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
            os.environ[k] = str(host_threads())
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
