#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/stage_bench.py --real > gpurun_out/s22_real.log 2>&1; echo "exit $?"; cat gpurun_out/s22_real.log | tail -12
