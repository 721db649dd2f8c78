#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SAD_B200_LIB=3dsad-main_b200/lib/libsad_prof.so timeout 120 python tools/sa_timeline.py sa1 0 2000 > gpurun_out/s13_tl_sa1.log 2>&1; echo "exit $?"
grep "gath" gpurun_out/s13_tl_sa1.log | sed -n 40,80p
