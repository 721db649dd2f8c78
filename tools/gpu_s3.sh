#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/stage_bench.py > gpurun_out/s3_bench.log 2>&1; echo "exit $?"; tail -1 gpurun_out/s3_bench.log
export SAD_B200_LIB=3dsad-main_b200/lib/libsad_prof.so
for s in sa4 sa1 sa2; do timeout 120 python tools/sa_timeline.py $s 0 400 > gpurun_out/s3_tl_$s.log 2>&1; echo "exit $?"; head -3 gpurun_out/s3_tl_$s.log; done
