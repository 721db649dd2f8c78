#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for s in sa1 sa1 sa2; do SAD_B200_LIB=3dsad-main_b200/lib/libsad_prof.so timeout 120 python tools/sa_timeline.py $s 300 260 > gpurun_out/s5_tl_$s.log 2>&1; echo "exit $?"; grep -A20 "FAILED\|==" gpurun_out/s5_tl_$s.log | head -24; done
