#!/bin/bash
mkdir -p gpurun_out
export DUPL_ORACLE_CACHE=/tmp/dupl_oracle_cache
LOG=gpurun_out/tests8.log; : > $LOG
for f in tests/test_gpu_augment.py tests/test_gpu_train.py tests/test_gpu_dense.py; do
echo "=== $f" >> $LOG; timeout 500 python -m pytest -q --no-header -p no:cacheprovider $f -m gpu -s >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|^E " $LOG | cut -c1-500 | tail -30
timeout 400 python bench.py --steps 10 --warmup 3 --phase C --no-cpu-baseline --no-reference-gpu --no-secondary --no-roofline > gpurun_out/bench8_C.json 2> gpurun_out/bench8_C.err; echo "bench C exit=$?"
grep '^{' gpurun_out/bench8_C.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print(round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s; e2e', round(d['e2e']['value'],1), 'loss', d['loss'])"
tail -3 gpurun_out/bench8_C.err
