#!/bin/bash
# ncu --set full of the throughput-policy FPS kernel (one launch, 8 x 40k -> 2048), after the same command ran plain
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/fps1_probe.py 8 > gpurun_out/ncu_fps_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fps_cull1_kernel" -c 2 -o gpurun_out/r02_fps -f python tools/fps1_probe.py 8 > gpurun_out/ncu_fps.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_fps.log; ls -la gpurun_out/r02_fps.ncu-rep
