"""bench.py contract on CPU: the reference arm (the oracle's C/OpenMP port on the host cores; the mounted reference has
no implementation) prints exactly ONE JSON line on stdout carrying the keys the driver reads, whatever else is written
to stderr."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "scenes/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert abs(d["e2e"]["value"] - d["value"]) < 1e-9
    assert "40k" in d["metric"] and "workload" in d["config"]


def test_our_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without CUDA the product arm exits with a message instead of timing something else."""
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]


def test_clock_sampler_window_selection():
    """The `clocks` block: median SM clock and throttle reasons of the samples inside the timed window; a window shorter
    than one polling period (the driver's 20-step run is ~7 ms) takes the nearest sample instead of reporting nothing."""
    import threading
    sys.path.insert(0, ROOT)
    import bench

    class NV:
        nvmlClocksEventReasonHwSlowdown, nvmlClocksEventReasonHwThermalSlowdown = 1, 2
        nvmlClocksEventReasonSwThermalSlowdown, nvmlClocksEventReasonSwPowerCap = 4, 8

    def sampler(rows):
        s = bench.ClockSampler.__new__(bench.ClockSampler)
        s.ok, s._stop, s.nv, s.smax, s.rows = True, False, NV, 1965.0, list(rows)
        s.th = threading.Thread(target=lambda: None)
        s.th.start()
        return s

    rows = [(1.0, 1900.0, 0), (2.0, 1965.0, 8), (3.0, 1950.0, 2)]
    inside = sampler(rows).stop(0.5, 2.5)
    assert inside == {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 2}
    short = sampler(rows).stop(2.8, 2.9)                       # no sample inside: the nearest one (t = 3.0)
    assert short["samples"] == 1 and short["sm_mhz"] == 1950.0 and short["reasons"] == ["hw_thermal_slowdown"]
    none = sampler([]).stop(0.0, 1.0)
    assert none["sm_mhz"] is None and none["samples"] == 0
