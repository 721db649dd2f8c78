#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s25_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s25_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s25_smoke.log 2>&1; echo "smoke exit $?"
timeout 600 python bench.py > gpurun_out/s25_bench.json 2> gpurun_out/s25_bench.err; echo "bench exit $?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/s25_bench20.json 2> gpurun_out/s25_bench20.err; echo "bench20 exit $?"
timeout 600 python bench.py --steps 2 --warmup 1 --no-hbm > gpurun_out/s25_b2.json 2>/dev/null && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s25_launches.csv python bench.py --steps 2 --warmup 1 --no-hbm > gpurun_out/s25_ncu.log 2>&1; echo "ncu exit $?"
