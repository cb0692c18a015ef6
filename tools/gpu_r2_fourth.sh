#!/bin/bash
mkdir -p gpurun_out
export DUPL_ORACLE_CACHE=/tmp/dupl_oracle_cache
LOG=gpurun_out/tests4.log; : > $LOG
for f in "tests/test_gpu_baseline_sizes.py -k attention" tests/test_gpu_train.py; do
  echo "=== $f" >> $LOG; timeout 900 python -m pytest -q --no-header -p no:cacheprovider $f -m gpu -s >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|^E |^attention" $LOG | cut -c1-400 | tail -20
TRAIN_ONLY=1 python tools/attn_bench.py
TRAIN_ONLY=1 REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_dq|attn_bwd_dkv" -s 6 -c 2 -o gpurun_out/prof_attn_bwd2 -f python tools/attn_bench.py > gpurun_out/ncu_attn_bwd2.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_attn_bwd2.ncu-rep 2>/dev/null | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench exit=$?"
grep '^{' gpurun_out/bench4.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print(round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s; e2e', round(d['e2e']['value'],1), 'roofline', round(d['roofline']['frac'],3))"
tail -3 gpurun_out/bench4.err
