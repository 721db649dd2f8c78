#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for s in sa1 sa2; do SAD_B200_LIB=3dsad-main_b200/lib/libsad_prof.so timeout 120 python tools/sa_timeline.py $s 0 3000 > gpurun_out/s24_tl_$s.log 2>&1; echo "exit $?"; done
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_modules_gpu.py -x -q -m gpu > gpurun_out/s24_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s24_tests.log
