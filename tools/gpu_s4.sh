#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_fast_gpu.py -x -q > gpurun_out/s4_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s4_tests.log
SAD_B200_LIB=3dsad-main_b200/lib/libsad_dbg.so timeout 300 python tools/stage_bench.py > gpurun_out/s4_bench_dbg.log 2>&1; echo "dbg exit $?"; grep -c stage gpurun_out/s4_bench_dbg.log; grep "timeout" gpurun_out/s4_bench_dbg.log | head -5
timeout 300 python tools/stage_bench.py > gpurun_out/s4_bench_def.log 2>&1; echo "def exit $?"; tail -1 gpurun_out/s4_bench_def.log
timeout 300 python tools/stage_bench.py > gpurun_out/s4_bench_def2.log 2>&1; echo "def2 exit $?"; grep -c stage gpurun_out/s4_bench_def2.log
for s in sa1 sa2 sa3; do SAD_B200_LIB=3dsad-main_b200/lib/libsad_prof.so timeout 120 python tools/sa_timeline.py $s 300 260 > gpurun_out/s4_tl_$s.log 2>&1; echo "exit $?"; done
