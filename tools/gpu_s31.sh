#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_tf32_gpu.py -x -q -m gpu > gpurun_out/s31_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s31_tests.log
SAD_B200_LIB=3dsad-main_b200/lib/libsad_tfprof.so timeout 120 python tools/tf32_probe.py > gpurun_out/s31_probe.log 2>&1; echo "probe exit $?"
timeout 120 python tools/tf32_probe.py > gpurun_out/s31_probe_plain.log 2>&1; echo "probe exit $?"
