#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/sub
run() { n=$1; shift
  timeout 300 "$@" > gpurun_out/sub/$n.json 2> gpurun_out/sub/$n.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/sub/$n.json"))
    print("$n", d["value"], d["e2e"]["value"], d["ms_per_step"], d["run"].get("host_submit_us_per_step"), d["run"].get("host_submit_us_first_calls")[:8], flush=True)
except Exception as e:
    print("$n failed", e); print(open("gpurun_out/sub/$n.err").read()[-1500:])
P
}
run res20 python bench.py --steps 20 --warmup 3 --no-hbm --no-cpu
run res20b python bench.py --steps 20 --warmup 3 --no-hbm --no-cpu
run res200 python bench.py --steps 200 --no-hbm --no-cpu
