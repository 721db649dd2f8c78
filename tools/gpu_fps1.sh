#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
L=3dsad-main_b200/lib
echo "== product"; timeout 120 python tools/fps1_probe.py 8
timeout 120 python tools/fps1_probe.py 8 throughput_paired
timeout 120 python tools/fps1_probe.py 1
echo "== profile build"; SAD_B200_LIB=$L/libsad_fpsprof.so timeout 120 python tools/fps1_probe.py 1 2>&1 | tail -8
for k in 1 2 3 4; do echo "== ablate $k"; SAD_B200_LIB=$L/libsad_fpsabl$k.so timeout 120 python tools/fps1_probe.py 8 2>&1 | tail -1; done
