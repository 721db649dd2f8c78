#!/bin/bash
# GPU tests, smoke() and the two contract lines (default and the driver's 20 steps) on the final tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/final
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "tests exit $?"; tail -3 $O/tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke exit $?"
timeout 300 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-hbm --no-cpu > $O/bench20.json 2> $O/bench20.err; echo "bench20 exit $?"
python - <<'P'
import json
for f in ("bench", "bench20"):
    d = json.load(open(f"gpurun_out/final/{f}.json"))
    print(f, d["value"], d["e2e"]["value"], d["ms_per_step"], d["run"]["host_submit_us_per_batch"], d["roofline"]["frac"], d["clocks"])
P
