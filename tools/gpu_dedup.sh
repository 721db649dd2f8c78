#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_fast_gpu.py -x -q -m gpu > gpurun_out/dd_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/dd_tests.log
timeout 300 python tools/dedup_probe.py 2>&1 | tail -8
