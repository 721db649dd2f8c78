#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/abl
L=3dsad-main_b200/lib
for v in b200 evl d8 evld8; do echo "== $v"; SAD_B200_LIB=$L/libsad_$v.so timeout 120 python tools/fps1_probe.py 8 2>&1 | tail -1; done
run() { n=$1; shift
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
for v in b200 evl evld8; do
  SAD_B200_LIB=$L/libsad_$v.so run ${v}_200 python bench.py --steps 200 --no-hbm --no-cpu
  SAD_B200_LIB=$L/libsad_$v.so run ${v}_20 python bench.py --steps 20 --warmup 3 --no-hbm --no-cpu
done
