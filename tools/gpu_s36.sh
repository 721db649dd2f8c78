#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/config_table.py --out gpurun_out/r02_configs.json > gpurun_out/s36_configs.log 2>&1; echo "configs exit $?"; tail -5 gpurun_out/s36_configs.log
timeout 300 python tools/stage_bench.py > gpurun_out/s36_stage_graph.log 2>&1; echo "stage exit $?"
timeout 900 python tools/sweep.py --out gpurun_out/r02_sweep.json > gpurun_out/s36_sweep.log 2>&1; echo "sweep exit $?"; tail -3 gpurun_out/s36_sweep.log
timeout 600 python tools/train_bench.py > gpurun_out/s36_train1.json 2> gpurun_out/s36_train1.err; echo "train exit $?"; tail -1 gpurun_out/s36_train1.json
