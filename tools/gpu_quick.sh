#!/bin/bash
# tests (per-file, isolated) + bench. Outputs under gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/tests.log; : > $LOG
for f in ${TEST_FILES:-tests/test_gpu_dense.py tests/test_gpu_cam_par.py tests/test_gpu_golden.py tests/test_gpu_crf.py}; do
  echo "=== $f" >> $LOG; timeout 600 python -m pytest -q -x --no-header -p no:cacheprovider $f -m gpu >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|assert" $LOG | tail -30
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
