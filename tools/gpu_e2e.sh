#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/e2e
run() { n=$1; shift
  timeout 300 python bench.py --no-hbm --no-cpu "$@" > gpurun_out/e2e/$n.json 2> gpurun_out/e2e/$n.err
  python - <<P
import json
d=json.load(open("gpurun_out/e2e/$n.json"))
print("$n", d["steps"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], flush=True)
P
}
for sl in 16 20 24 28; do
  run s200_slots$sl --slots $sl --copy-inputs
  run s20_slots$sl --slots $sl --steps 20 --warmup 3 --copy-inputs
done
