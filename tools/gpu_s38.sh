#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/train_bench.py > gpurun_out/s38_train.json 2> gpurun_out/s38_train.err; echo "train exit $?"; tail -1 gpurun_out/s38_train.json
timeout 600 python tools/train_bench.py --deterministic > gpurun_out/s38_train_det.json 2> gpurun_out/s38_train_det.err; echo "train det exit $?"; tail -1 gpurun_out/s38_train_det.json
timeout 600 python tools/train_bench.py --amp > gpurun_out/s38_train_amp.json 2> gpurun_out/s38_train_amp.err; echo "train amp exit $?"; tail -1 gpurun_out/s38_train_amp.json
