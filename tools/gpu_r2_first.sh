#!/bin/bash
# Round 2, first GPU call: the new BASELINE-size parity tests first, then the rest of the GPU suite, the default bench line
# (train step + secondaries + reference_gpu), a short reference arm, and the precision table.
mkdir -p gpurun_out
export DUPL_ORACLE_CACHE=/tmp/dupl_oracle_cache
LOG=gpurun_out/tests.log; : > $LOG
nproc >> $LOG
for f in tests/test_gpu_baseline_sizes.py tests/test_gpu_train.py tests/test_gpu_m1_script.py ${MORE_TESTS}; do
  echo "=== $f" >> $LOG; timeout ${TEST_TIMEOUT:-1500} python -m pytest -q --no-header -p no:cacheprovider $f -m gpu -s >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|^E |^attention|^mscam|^train448|^coco" $LOG | cut -c1-600 | tail -80
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit=$?"
tail -c 6000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref exit=$?"
tail -c 3000 gpurun_out/bench_reference.json; tail -5 gpurun_out/bench_reference.err
timeout 1500 python tools/precision_table.py > gpurun_out/precision_table.log 2>&1; echo "precision exit=$?"
tail -25 gpurun_out/precision_table.log
