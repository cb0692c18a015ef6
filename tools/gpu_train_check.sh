#!/bin/bash
# GEMM + training tests, then the training-step bench (A/B: fixed tiling / no forward reuse).  Outputs: gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/tests.log; : > $LOG
for f in ${TEST_FILES:-tests/test_gpu_dense.py tests/test_gpu_train.py}; do
  echo "=== $f" >> $LOG; timeout ${TEST_TIMEOUT:-400} python -m pytest -q -x --no-header -p no:cacheprovider $f -m gpu >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|assert" $LOG | tail -40
for mode in default ${AB_MODES}; do
  case $mode in
    default) env_args="";;
    fixed) env_args="DUPL_GEMM_TILING=fixed";;
    noreuse) env_args="DUPL_NO_REUSE=1";;
    *) env_args="$mode";;
  esac
  env $env_args timeout 400 python tools/bench_train.py --steps 8 --warmup 3 ${PROFILE_FLAG} > gpurun_out/bench_train_$mode.json 2> gpurun_out/bench_train_$mode.err; echo "train[$mode] exit=$?"
  python -c "
import json
d=json.load(open('gpurun_out/bench_train_$mode.json'))
print('$mode', 'img/s', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'loss', d['loss_last'], 'mem', round(d['peak_mem_gb'],1))"
done
