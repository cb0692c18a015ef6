#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/tests.log; : > $LOG
for f in ${TEST_FILES:-tests/test_gpu_train.py tests/test_gpu_eval_sweep.py tests/test_gpu_dense.py}; do
  echo "=== $f" >> $LOG; timeout ${TEST_TIMEOUT:-500} python -m pytest -q -x --no-header -p no:cacheprovider $f -m gpu >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|assert|^E " $LOG | tail -60
for w in ${WORKLOADS:-cam_par train crf_sweep}; do
  timeout 600 python bench.py --workload $w --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench[$w] exit=$?"
  cat gpurun_out/bench_$w.json; tail -4 gpurun_out/bench_$w.err
done
