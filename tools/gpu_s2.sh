#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/stage_bench.py > gpurun_out/s2_bench.log 2>&1; echo "exit $?"; tail -3 gpurun_out/s2_bench.log
timeout 300 python tools/stage_bench.py --tpc 6 > gpurun_out/s2_bench6.log 2>&1; echo "exit $?"; tail -1 gpurun_out/s2_bench6.log
timeout 600 python -m pytest tests/test_mlp_gpu.py tests/test_modules_gpu.py tests/test_engine_gpu.py -x -q > gpurun_out/s2_tests.log 2>&1; echo "exit $?"; tail -5 gpurun_out/s2_tests.log
