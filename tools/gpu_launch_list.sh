#!/bin/bash
# ncu launch list of bench.py (only after the same command has exited 0 without ncu); summarise with tools/launch_list_summary.py
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 6 --warmup 3 --slots 2 --no-hbm --no-cpu > gpurun_out/launch_list_plain.json 2>/dev/null && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 --slots 2 --no-hbm --no-cpu > gpurun_out/launch_list_ncu.log 2>&1; echo "ncu exit $?"
