#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --mlp-dtype tf32 --no-hbm --no-cpu > gpurun_out/s28_bench_tf32.json 2> gpurun_out/s28_bench_tf32.err; echo "tf32 bench exit $?"; tail -3 gpurun_out/s28_bench_tf32.err
timeout 600 python bench.py > gpurun_out/s28_bench.json 2> gpurun_out/s28_bench.err; echo "bench exit $?"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s28_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s28_tests.log
