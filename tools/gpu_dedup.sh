#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/dd_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/dd_tests.log
timeout 600 python bench.py --no-cpu > gpurun_out/dd_bench.json 2> gpurun_out/dd_bench.err; echo "bench exit $?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-hbm > gpurun_out/dd_bench20.json 2> gpurun_out/dd_bench20.err; echo "bench20 exit $?"
