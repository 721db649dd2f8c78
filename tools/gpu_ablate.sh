#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/abl
run() { # name, cmd...
  n=$1; shift
  timeout 300 "$@" > gpurun_out/abl/$n.json 2> gpurun_out/abl/$n.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/abl/$n.json"))
    print("$n", d["value"], d["e2e"]["value"], d["ms_per_step"], d["run"].get("batch_latency_ms"), flush=True)
except Exception as e:
    print("$n failed", e); print(open("gpurun_out/abl/$n.err").read()[-800:])
P
}
run base200 python bench.py --steps 200 --no-hbm --no-cpu
run nofps200 python tools/ablate_step.py nofps --steps 200 --no-hbm --no-cpu
run latency200 python bench.py --steps 200 --no-hbm --no-cpu --fps-policy latency
run nofps20 python tools/ablate_step.py nofps --steps 20 --no-hbm --no-cpu
