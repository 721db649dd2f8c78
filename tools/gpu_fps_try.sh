#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/spec
timeout 100 python tools/fps1_probe.py 8 2>&1 | grep variant
timeout 300 python -m pytest tests/test_grid_gpu.py -x -q -m gpu -k "fps" > gpurun_out/spec/tests.log 2>&1; echo "fps tests exit $?"; tail -3 gpurun_out/spec/tests.log
timeout 300 python bench.py --no-hbm --no-cpu > gpurun_out/spec/b200.json 2>/dev/null
python - <<P
import json
d=json.load(open("gpurun_out/spec/b200.json")); print(200, d["value"], d["e2e"]["value"], d["ms_per_step"], d["run"]["batch_latency_ms"])
P
