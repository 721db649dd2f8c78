#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/s17_tests.log 2>&1; echo "tests exit $?"; tail -6 gpurun_out/s17_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s17_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/s17_smoke.log
timeout 900 python bench.py > gpurun_out/s17_bench.json 2> gpurun_out/s17_bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/s17_bench.json; tail -3 gpurun_out/s17_bench.err
timeout 600 python bench.py --no-cpu --no-hbm --steps 20 > gpurun_out/s17_bench20.json 2> gpurun_out/s17_bench20.err; echo "bench20 exit $?"; cut -c1-200 gpurun_out/s17_bench20.json
