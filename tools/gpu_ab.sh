#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/tests.log; : > $LOG
for f in tests/test_gpu_dense.py tests/test_gpu_golden.py; do
  echo "=== $f (BK=32)" >> $LOG; timeout 600 python -m pytest -q -x --no-header -p no:cacheprovider $f -m gpu >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|assert" $LOG | tail -30
for bk in 32 64; do
  DUPL_GEMM_BK=$bk timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/bench_bk$bk.json 2>> gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_bk$bk.json'))
print('BK=$bk', 'img/s', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved'],1), d['breakdown_ms'])
PY
done
DUPL_GEMM_BK=32 timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline --fuse-students > gpurun_out/bench_fused.json 2>> gpurun_out/bench.err
python -c "
import json
d=json.load(open('gpurun_out/bench_fused.json'))
print('fused BK=32', 'img/s', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved'],1), d['breakdown_ms'])"
tail -3 gpurun_out/bench.err
