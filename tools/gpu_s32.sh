#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SAD_B200_LIB=3dsad-main_b200/lib/libsad_tfprof.so timeout 120 python tools/tf32_probe.py sa1 sa2 fp1 > gpurun_out/s32_probe.log 2>&1; echo "probe exit $?"
timeout 600 python -m pytest tests/test_mlp_tf32_gpu.py -x -q -m gpu > gpurun_out/s32_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s32_tests.log
