#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SAD_B200_LIB=3dsad-main_b200/lib/libsad_dbg.so timeout 300 python tools/stage_bench.py > gpurun_out/s4_bench_dbg.log 2>&1; echo "dbg exit $?"; grep -c stage gpurun_out/s4_bench_dbg.log; grep "timeout" gpurun_out/s4_bench_dbg.log | head -5
timeout 300 python tools/stage_bench.py > gpurun_out/s4_bench_def.log 2>&1; echo "def exit $?"; grep -c stage gpurun_out/s4_bench_def.log
timeout 300 python tools/stage_bench.py > gpurun_out/s4_bench_def2.log 2>&1; echo "def2 exit $?"; grep -c stage gpurun_out/s4_bench_def2.log
SAD_B200_LIB=3dsad-main_b200/lib/libsad_prof.so timeout 120 python tools/sa_timeline.py sa1 600 300 > gpurun_out/s4_tl_sa1.log 2>&1; echo "exit $?"
