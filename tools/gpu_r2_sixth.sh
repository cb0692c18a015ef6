#!/bin/bash
mkdir -p gpurun_out
export DUPL_ORACLE_CACHE=/tmp/dupl_oracle_cache
LOG=gpurun_out/tests6.log; : > $LOG
for f in tests/test_gpu_optim.py tests/test_gpu_train.py; do
echo "=== $f" >> $LOG; timeout 400 python -m pytest -q --no-header -p no:cacheprovider $f -m gpu -s >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|^E " $LOG | cut -c1-400 | tail -20
for fused in 1 0; do
DUPL_FUSED_ADAMW=$fused timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary > gpurun_out/bench6_$fused.json 2> gpurun_out/bench6_$fused.err; echo "bench fused=$fused exit=$?"
grep '^{' gpurun_out/bench6_$fused.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print(round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s; e2e', round(d['e2e']['value'],1), 'roofline', round(d['roofline']['frac'],3), 'loss', d['loss'])"
tail -3 gpurun_out/bench6_$fused.err
done
