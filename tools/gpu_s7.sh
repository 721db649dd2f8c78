#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_fast_gpu.py -x -q > gpurun_out/s8_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s8_tests.log
for i in 1 2; do timeout 300 python tools/stage_bench.py > gpurun_out/s8_bench$i.log 2>&1; echo "bench$i exit $?"; done; tail -1 gpurun_out/s8_bench1.log
timeout 300 python tools/stage_bench.py --tpc 6 > gpurun_out/s8_bench_tpc6.log 2>&1; echo "tpc6 exit $?"
for s in sa1 sa2 sa3 sa4; do SAD_B200_LIB=3dsad-main_b200/lib/libsad_prof.so timeout 120 python tools/sa_timeline.py $s 0 500 > gpurun_out/s8_tl_$s.log 2>&1; echo "exit $?"; done
