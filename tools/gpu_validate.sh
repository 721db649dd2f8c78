#!/bin/bash
# Full single-GPU validation: GPU tests, smoke(), the bench lines kept under profiles/ (default, the driver's 20 steps,
# tf32, CPU reference arm) and two scheduling variants.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/final
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "tests exit $?"; tail -3 $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke exit $?"
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench20.json 2> $O/bench20.err; echo "bench20 exit $?"
timeout 600 python bench.py --mlp-dtype tf32 --no-hbm --no-cpu > $O/bench_tf32.json 2> $O/bench_tf32.err; echo "tf32 exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref exit $?"
for v in "paired200 --fps-policy throughput_paired" "paired20 --fps-policy throughput_paired --steps 20 --warmup 3" "slots48 --slots 48" "py_forked_copy20 --py-submit --forked-graph --copy-inputs --steps 20 --warmup 3"; do
  set -- $v; n=$1; shift
  timeout 300 python bench.py --no-hbm --no-cpu "$@" > $O/$n.json 2> $O/$n.err
done
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/final/*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], d.get("value"), (d.get("e2e") or {}).get("value"), d.get("ms_per_step"), (d.get("run") or {}).get("host_submit_us_per_step"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
P
