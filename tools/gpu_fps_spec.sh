#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/spec
for d in 3 5 6; do echo "DEPTH=$d"; SAD_B200_LIB=3dsad-main_b200/lib/libsad_d$d.so timeout 100 python tools/fps1_probe.py 8 2>&1 | grep "variant 0"; done
for t in 3 4 8 12; do
timeout 300 python bench.py --no-hbm --no-cpu --tpc $t > gpurun_out/spec/tpc$t.json 2>/dev/null
python - <<P
import json
d=json.load(open("gpurun_out/spec/tpc$t.json")); print("tpc $t", d["value"], d["e2e"]["value"], d["ms_per_step"])
P
done
