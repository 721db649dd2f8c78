"""Summarise the ncu launch list of `bench.py` (gpu__time_duration.sum, --clock-control none) into per-kernel shares of
ONE timed step.

    python tools/launch_list_summary.py gpurun_out/launches.csv profiles/rNN_launches_summary.json

A step (one batch through the pipelined executor = one CUDA-graph replay) starts at `grid_build_kernel`; the graph
replays all have the same launch count, so the most frequent segment length is a replayed step and the LAST segment of
that length is the one summarised.  ncu times are cold-cache (L2 flushed per kernel) and serialised: compare SHARES with
bench.py's live CUDA-event numbers, not absolutes."""
import collections
import csv
import json
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    starts = [i for i, r in enumerate(rows) if "grid_build_kernel" in r["Kernel Name"]]
    segs = [rows[a:b] for a, b in zip(starts, starts[1:] + [len(rows)])]
    lens = collections.Counter(len(s) for s in segs[:-1])
    per = lens.most_common(1)[0][0]
    step = [s for s in segs[:-1] if len(s) == per][-1]
    agg, tot = collections.OrderedDict(), 0.0
    for row in step:
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = v / 1000 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1000)
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "").strip()[-70:]
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += us
        tot += us
    kernels = [{"kernel": k, "launches": c, "us": round(us, 1), "share": round(us / tot, 4)}
               for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    lib = sum(k["launches"] for k in kernels if not k["kernel"].startswith("at::"))
    out = {"source": src, "launches_total": len(rows), "segments": len(segs), "segments_of_this_length": lens[per],
           "launches_in_step": per, "library_launches_in_step": lib, "torch_launches_in_step": per - lib,
           "step_us_ncu": round(tot, 1),
           "note": "ncu per-launch times are cold-cache and serialised; shares are what to compare", "kernels": kernels}
    json.dump(out, open(dst, "w"), indent=1)
    for k in kernels:
        print(f"{k['us']:9.1f} us {100 * k['share']:5.1f}%  x{k['launches']:3d}  {k['kernel']}")
    print("step total (ncu) us:", round(tot, 1), "launches:", per, "torch:", per - lib, "replayed steps seen:", lens[per])


if __name__ == "__main__":
    main()
