#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_tf32_gpu.py -x -q -m gpu > gpurun_out/tf_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/tf_tests.log
timeout 120 python tools/tf32_probe.py 2>&1 | grep us
timeout 600 python bench.py --mlp-dtype tf32 --no-hbm --no-cpu > gpurun_out/tf_bench.json 2>/dev/null; echo "bench exit $?"
