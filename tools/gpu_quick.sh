#!/bin/bash
# tests (per-file, isolated) + bench. Outputs under gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/tests.log; : > $LOG
for f in ${TEST_FILES:-tests/test_gpu_dense.py tests/test_gpu_cam_par.py tests/test_gpu_golden.py tests/test_gpu_crf.py tests/test_gpu_losses.py}; do
  echo "=== $f" >> $LOG; timeout ${TEST_TIMEOUT:-150} python -m pytest -q -x --no-header -p no:cacheprovider $f -m gpu >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|assert" $LOG | tail -30
timeout ${BENCH_TIMEOUT:-240} python bench.py --steps 10 --warmup 3 --breakdown ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench.json'))
print('img/s', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'gemm TF', round(d['roofline']['achieved'],1), 'frac', round(d['roofline']['frac'],3), d['clocks'], d['breakdown_ms'])"
tail -5 gpurun_out/bench.err
