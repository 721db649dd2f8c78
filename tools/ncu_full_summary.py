"""Summarise `ncu --set full` reports into profiles/ncu_full_summary.json (what bench.py's
roofline.traffic reads) -- one row per profiled launch, keyed the way bench.py names kernels.

    python tools/ncu_full_summary.py gpurun_out/r01_mlp.ncu-rep gpurun_out/r01_search.ncu-rep ... \
        --out profiles/ncu_full_summary.json

Reads the reports here (no GPU needed): `ncu -i <rep> --page raw --csv`."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_insts",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_insts",
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6,
              "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def key_of(name, grid, block):
    n = name
    if "fused_mlp" in n:
        return "shared_mlp"
    if "sa_mlp_kernel" in n:
        return "sa_mlp"
    if "pw_mlp_kernel" in n:
        return "pw_mlp"
    if "mlp_tf32" in n:
        return "mlp_tf32"
    if "interp_fwd" in n:
        return "three_interpolate"
    if "fps_cull" in n:
        return "furthest_point_sample_grid"
    if "fps_kernel" in n:
        return "furthest_point_sample"
    if "ball_query_grid" in n:
        return "ball_query_grid"
    if "ball_query" in n:
        return "ball_query"
    if "grid_build" in n:
        return "scene_grid_build"
    if "interp_cl" in n:
        return "three_interpolate_cl"
    if "three_nn" in n:
        return "three_nn"
    if "group_fwd" in n:
        return "grouping_operation"
    if "cf_to_cl" in n:
        return "cf_to_cl_bf16"
    return n[:40]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    col = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rd[2:]:
        if len(r) < len(hdr):
            continue
        d = {"kernel_name": r[col["Kernel Name"]][:90], "report": rep.split("/")[-1]}
        for m, k in WANT.items():
            if m in col:
                try:
                    v = float(r[col[m]].replace(",", ""))
                except ValueError:
                    continue
                d[k] = v * UNIT_SCALE.get(units[col[m]], 1)
        d["grid"] = r[col["Grid Size"]] if "Grid Size" in col else None
        d["block"] = r[col["Block Size"]] if "Block Size" in col else None
        d["roofline_key"] = key_of(d["kernel_name"], d["grid"], d["block"])
        if "dram_read" in d and "dram_write" in d:
            d["dram_bytes_per_launch"] = int(d["dram_read"] + d["dram_write"])
        res.append(d)
    return res


def main():
    reps = [a for a in sys.argv[1:] if a.endswith(".ncu-rep")]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else "profiles/ncu_full_summary.json"
    rows = []
    for rep in reps:
        rows += rows_of(rep)
    json.dump({"note": "one row per profiled launch (tools/ncu_target.py: the fused-MLP kernels at the benchmark's shapes, "
                       "B = 8 x 40k surface scenes; tf32 SA2; three_interpolate at the FP2 shape, B = 256); ncu --set full "
                       "--clock-control none; times are cold-cache and serialised",
               "kernels": rows}, open(out, "w"), indent=1)
    for d in rows:
        print(f"{d['roofline_key']:28s} grid {str(d['grid']):14s} {d.get('duration_us', 0):9.1f} us  dram "
              f"{d.get('dram_bytes_per_launch', 0) / 1e6:8.2f} MB  tensor {d.get('tensor_pipe_active_pct', 0):5.1f}%  "
              f"occ {d.get('achieved_occupancy_pct', 0):5.1f}%")


if __name__ == "__main__":
    main()
