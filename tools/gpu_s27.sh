#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_tf32_gpu.py -x -q -m gpu > gpurun_out/s27_tf32.log 2>&1; echo "tf32 exit $?"; tail -30 gpurun_out/s27_tf32.log
