#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
SAD_B200_LIB=3dsad-main_b200/lib/libsad_fpsprof.so timeout 100 python tools/fps1_probe.py 1 2>&1 | grep -E "fps_spec1|variant" | sort | uniq | head -12
