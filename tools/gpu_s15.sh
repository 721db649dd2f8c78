#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/stage_bench.py --no-cf > gpurun_out/s15_bench_nocf.log 2>&1; echo "exit $?"; grep "fast" gpurun_out/s15_bench_nocf.log | cut -c1-120
