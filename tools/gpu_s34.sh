#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SAD_B200_LIB=3dsad-main_b200/lib/libsad_tfspin.so timeout 120 python tools/tf32_probe.py sa1 sa2 fp1 > gpurun_out/s34_probe_spin.log 2>&1; echo "probe exit $?"
timeout 120 python tools/tf32_probe.py sa1 sa2 fp1 > gpurun_out/s34_probe.log 2>&1; echo "probe exit $?"
