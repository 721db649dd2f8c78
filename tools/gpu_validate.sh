#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s35_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s35_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s35_smoke.log 2>&1; echo "smoke exit $?"
timeout 600 python bench.py > gpurun_out/s35_bench.json 2> gpurun_out/s35_bench.err; echo "bench exit $?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/s35_bench20.json 2> gpurun_out/s35_bench20.err; echo "bench20 exit $?"
timeout 600 python bench.py --mlp-dtype tf32 --no-hbm --no-cpu > gpurun_out/s35_bench_tf32.json 2> gpurun_out/s35_bench_tf32.err; echo "tf32 exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s35_bench_ref.json 2> gpurun_out/s35_bench_ref.err; echo "ref exit $?"
bash tools/gpu_ncu.sh
