"""Summarise an ncu launch list (gpu__time_duration.sum CSV of tools/profile_step.py) into
per-kernel shares of ONE step.

    python tools/ncu_summary.py gpurun_out/launches.csv profiles/rNN_launches_summary.json [--passes 3] [--tail 2]

profile_step.py runs (warmup + steps) identical passes followed by `tail` trailing launches
(the checksum reduction); the last pass is the one summarised.  ncu times are cold-cache and
serialised: compare SHARES with bench.py's live CUDA-event numbers, not absolutes."""
import collections
import csv
import json
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    passes = int(sys.argv[sys.argv.index("--passes") + 1]) if "--passes" in sys.argv else 3
    tail = int(sys.argv[sys.argv.index("--tail") + 1]) if "--tail" in sys.argv else 2
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    # a pass starts at each launch of the step's first kernel (the first pass also carries
    # one-time weight folding / packing launches, so passes are not equally long)
    first = rows[0]["Kernel Name"]
    starts = [i for i, r in enumerate(rows) if r["Kernel Name"] == first]
    assert len(starts) == passes, f"expected {passes} passes, found {len(starts)}"
    step = rows[starts[-1]: len(rows) - tail]
    per = len(step)
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
    out = {"source": src, "launches_total": len(rows), "launches_in_step": per, "step_us_ncu": round(tot, 1),
           "note": "ncu per-launch times are cold-cache and serialised; shares are what to compare",
           "kernels": kernels}
    json.dump(out, open(dst, "w"), indent=1)
    for k in kernels[:14]:
        print(f"{k['us']:9.1f} us {100 * k['share']:5.1f}%  x{k['launches']:3d}  {k['kernel']}")
    print("step total (ncu) us:", round(tot, 1), "launches:", per)


if __name__ == "__main__":
    main()
