#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/stage_bench.py --real > gpurun_out/s26_real_graph.log 2>&1; echo "graph exit $?"
timeout 300 python tools/stage_bench.py --real --events > gpurun_out/s26_real_events.log 2>&1; echo "events exit $?"
timeout 300 python tools/stage_bench.py --events > gpurun_out/s26_events.log 2>&1; echo "events exit $?"
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_train_gpu.py tests/test_mlp_gpu.py -x -q -m gpu > gpurun_out/s26_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s26_tests.log
timeout 600 python bench.py --steps 6 --warmup 3 --slots 2 --no-hbm --no-cpu > gpurun_out/s26_b.json 2>/dev/null && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/s26_launches.csv python bench.py --steps 6 --warmup 3 --slots 2 --no-hbm --no-cpu > gpurun_out/s26_ncu.log 2>&1; echo "ncu exit $?"
