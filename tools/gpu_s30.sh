#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SAD_B200_LIB=3dsad-main_b200/lib/libsad_tfprof.so timeout 120 python tools/tf32_probe.py > gpurun_out/s30_probe.log 2>&1; echo "probe exit $?"
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k interpolate > gpurun_out/s30_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s30_tests.log
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu > gpurun_out/s30_bench.json 2> gpurun_out/s30_bench.err; echo "bench exit $?"
