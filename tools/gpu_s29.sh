#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_mlp_tf32_gpu.py -x -q -m gpu > gpurun_out/s29_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s29_tests.log
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu > gpurun_out/s29_bench.json 2> gpurun_out/s29_bench.err; echo "bench exit $?"
timeout 600 python bench.py --mlp-dtype tf32 --no-hbm --no-cpu > gpurun_out/s29_bench_tf32.json 2> gpurun_out/s29_bench_tf32.err; echo "tf32 bench exit $?"
