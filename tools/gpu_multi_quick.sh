#!/bin/bash
# usage: gpu_multi_quick.sh N [extra bench args] -- one bench.py run on N GPUs of one box (torchrun)
N=$1; shift
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --no-hbm --no-cpu "$@" > gpurun_out/r02_bench_n${N}_quick.json 2> gpurun_out/r02_bench_n${N}_quick.err
echo "n=$N exit $?"
python - <<P
import json
d=json.load(open("gpurun_out/r02_bench_n${N}_quick.json"))
print(d["n_gpus"], d["steps"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"])
P
