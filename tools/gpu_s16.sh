#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s16_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/s16_tests.log
timeout 600 python bench.py --no-cpu > gpurun_out/s16_bench.json 2> gpurun_out/s16_bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/s16_bench.json
timeout 600 python bench.py --no-cpu --steps 20 > gpurun_out/s16_bench20.json 2> gpurun_out/s16_bench20.err; echo "bench exit $?"; cut -c1-300 gpurun_out/s16_bench20.json
