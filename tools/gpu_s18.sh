#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/s18_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/s18_tests.log
for sl in 16 20 32; do timeout 600 python bench.py --no-cpu --no-hbm --steps 20 --slots $sl > gpurun_out/s18_b20_s$sl.json 2> /dev/null; echo "slots $sl: $(cut -c60-110 gpurun_out/s18_b20_s$sl.json)"; done
for sl in 20 32; do timeout 600 python bench.py --fps-policy throughput_paired --no-cpu --no-hbm --steps 20 --slots $sl > gpurun_out/s18_b20_sc2_s$sl.json 2> /dev/null; echo "paired slots $sl: $(cut -c60-110 gpurun_out/s18_b20_sc2_s$sl.json)"; done
timeout 600 python bench.py --fps-policy throughput_paired --no-cpu --no-hbm > gpurun_out/s18_b200_sc2.json 2> /dev/null; echo "paired 200: $(cut -c60-110 gpurun_out/s18_b200_sc2.json)"
timeout 600 python bench.py --no-cpu --no-hbm > gpurun_out/s18_b200.json 2> gpurun_out/s18_b200.err; echo "200: $(cut -c60-110 gpurun_out/s18_b200.json)"
