#!/bin/bash
# CUDA_DEVICE_MAX_CONNECTIONS sweep for the pipelined executor (32 slots = 32 streams; the default is 8 hardware queues)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/conn
for conn in 8 32; do
  for steps in 20 200; do
    for pol in throughput throughput_paired; do
      CUDA_DEVICE_MAX_CONNECTIONS=$conn timeout 300 python bench.py --steps $steps --warmup 3 --fps-policy $pol --no-hbm --no-cpu \
        > gpurun_out/conn/c${conn}_s${steps}_${pol}.json 2>gpurun_out/conn/c${conn}_s${steps}_${pol}.err
      python - <<P
import json
d=json.load(open("gpurun_out/conn/c${conn}_s${steps}_${pol}.json"))
print("conn $conn steps $steps $pol", d["value"], d["e2e"]["value"], d["ms_per_step"], flush=True)
P
    done
  done
done
