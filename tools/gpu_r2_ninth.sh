#!/bin/bash
mkdir -p gpurun_out
export DUPL_ORACLE_CACHE=/tmp/dupl_oracle_cache
LOG=gpurun_out/tests9.log; : > $LOG
for f in tests/test_gpu_train.py "tests/test_gpu_baseline_sizes.py -k phase_b"; do
echo "=== $f" >> $LOG; timeout 500 python -m pytest -q --no-header -p no:cacheprovider $f -m gpu -s >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|^E |^train448" $LOG | cut -c1-400 | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "bench exit=$?"
grep '^{' gpurun_out/bench9.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print(round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s; e2e', round(d['e2e']['value'],1), 'roofline', round(d['roofline']['frac'],3), 'loss', d['loss'], 'launches', d['gpu_launches'])"
tail -3 gpurun_out/bench9.err
timeout 200 python tools/ncu_step.py --table > gpurun_out/r02_train_step_phaseB_kernels.txt 2> gpurun_out/ncu_table.err; tail -2 gpurun_out/r02_train_step_phaseB_kernels.txt | cut -c1-120
