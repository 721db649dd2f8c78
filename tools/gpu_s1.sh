#!/bin/bash
# bring-up of the specialised SA kernel: instance by instance, each under its own timeout
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 240 "$@" > gpurun_out/s1_$name.log 2>&1; echo "exit $?"; tail -5 gpurun_out/s1_$name.log; }
run sa1 python -m pytest tests/test_mlp_fast_gpu.py -x -q -k "sa1"
run sa2single python -m pytest tests/test_mlp_fast_gpu.py -x -q -k "sa2-single"
run sa2pair python -m pytest tests/test_mlp_fast_gpu.py -x -q -k "sa2-pair"
run sa3pair python -m pytest tests/test_mlp_fast_gpu.py -x -q -k "sa3-pair or agg"
run bench python tools/stage_bench.py
