#!/bin/bash
# usage: gpu_multi.sh N  -- bench.py on N GPUs of one box (torchrun), default + the driver's 20-step window
N=$1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { # tag, extra args
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N "$@" > gpurun_out/r02_bench_n${N}${tag}.json 2> gpurun_out/r02_bench_n${N}${tag}.err
  echo "n=$N $tag exit $?"; tail -c 400 gpurun_out/r02_bench_n${N}${tag}.json | head -c 400; echo
}
run "" --no-hbm
run "_20steps" --steps 20 --warmup 3 --no-hbm
if [ "$N" = "8" ]; then
  run "_20k" --points 20000 --no-hbm
  run "_200k" --points 200000 --steps 60 --slots 8 --sets 4 --no-hbm
fi
