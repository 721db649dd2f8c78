#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/s19_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/s19_tests.log
timeout 600 python bench.py --no-cpu --no-hbm --steps 20 > gpurun_out/s19_b20.json 2> gpurun_out/s19_b20.err; echo "20: $(cut -c60-110 gpurun_out/s19_b20.json)"; tail -2 gpurun_out/s19_b20.err
timeout 600 python bench.py --no-cpu --no-hbm > gpurun_out/s19_b200.json 2> gpurun_out/s19_b200.err; echo "200: $(cut -c60-110 gpurun_out/s19_b200.json)"
