#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/s23_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/s23_tests.log
timeout 300 python tools/stage_bench.py > gpurun_out/s23_bench.log 2>&1; echo "bench exit $?"; grep "fast\|vote" gpurun_out/s23_bench.log | grep -v "^{\"tpc" | cut -c1-150
timeout 300 python tools/stage_bench.py --real > gpurun_out/s23_real.log 2>&1; tail -6 gpurun_out/s23_real.log
timeout 600 python bench.py --no-cpu --no-hbm --steps 20 > gpurun_out/s23_b20.json 2> gpurun_out/s23_b20.err; echo "20: $(cut -c60-110 gpurun_out/s23_b20.json)"; tail -2 gpurun_out/s23_b20.err
timeout 600 python bench.py --no-cpu --no-hbm > gpurun_out/s23_b200.json 2> gpurun_out/s23_b200.err; echo "200: $(cut -c60-110 gpurun_out/s23_b200.json)"
