#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for nw in 20 24 28 32; do echo "NW=$nw"; SAD_FPS1_NW=$nw SAD_B200_LIB=3dsad-main_b200/lib/libsad_tools.so timeout 100 python tools/fps1_probe.py 8 2>&1 | grep "variant 0"; done
