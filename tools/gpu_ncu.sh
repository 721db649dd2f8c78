#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/ncu_target.py 2 > gpurun_out/ncu_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"sa_mlp_kernel|pw_mlp_kernel|mlp_tf32_kernel|interp_fwd_pipe" -c 14 -o gpurun_out/r02_mlp -f python tools/ncu_target.py 2 > gpurun_out/ncu_mlp.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_mlp.log; ls -la gpurun_out/r02_mlp.ncu-rep
