#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for tpc in 1 2 3 4 6 8; do
  timeout 300 python bench.py --tpc $tpc --steps 20 --warmup 3 --no-cpu --no-hbm > gpurun_out/exp20_tpc$tpc.json 2>/dev/null; echo "$tpc exit $?"
  timeout 300 python bench.py --tpc $tpc --steps 20 --warmup 3 --no-cpu --no-hbm > gpurun_out/exp20b_tpc$tpc.json 2>/dev/null
  timeout 300 python bench.py --tpc $tpc --no-cpu --no-hbm > gpurun_out/exp_tpc$tpc.json 2>/dev/null
done
