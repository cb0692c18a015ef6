#!/bin/bash
# Round checkpoint on the B200 box: full GPU test suite (one process per file), smoke(), then bench.py for the three
# workloads.  Outputs: gpurun_out/.   TEST_FILES / WORKLOADS / STEPS narrow the run.
mkdir -p gpurun_out
LOG=gpurun_out/tests.log; : > $LOG
for f in ${TEST_FILES:-tests/test_gpu_*.py}; do
  echo "=== $f" >> $LOG; timeout ${TEST_TIMEOUT:-400} python -m pytest -q -x --no-header -p no:cacheprovider $f -m gpu >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|assert|^E " $LOG | tail -60
if [ -z "$NO_SMOKE" ]; then timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; fi
for w in ${WORKLOADS:-cam_par train crf_sweep}; do
  timeout 600 python bench.py --workload $w --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench[$w] exit=$?"
  grep '^{' gpurun_out/bench_$w.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$w', round(d['value'], 2), d['unit'], round(d['ms_per_step'], 2), 'ms/step; e2e', round(d['e2e']['value'], 2), '; roofline', d['roofline']['kernel'][:30], round(d['roofline']['frac'], 3), '; train_step', (d.get('train_step') or {}).get('ms_per_step'), d.get('breakdown_ms', ''))"
  tail -3 gpurun_out/bench_$w.err
done
