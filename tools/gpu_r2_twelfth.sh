#!/bin/bash
# validate: column sums finished inside split_transpose (tickets); A/B: separate finish kernel, shared-memory carve-out; gap attribution
mkdir -p gpurun_out
LOG=gpurun_out/tests12.log; : > $LOG
for f in tests/test_gpu_train.py; do
echo "=== $f" >> $LOG; timeout 400 python -m pytest -q --no-header -p no:cacheprovider $f -m gpu >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|^E |^FAILED" $LOG | cut -c1-300 | tail -20
summ() { grep '^{' $1 | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('$1', round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s; e2e', round(d['e2e']['value'],1), 'loss', d.get('loss'), 'launches', d.get('gpu_launches'), 'clk', d.get('clocks',{}).get('sm_mhz'))
"; }
B="--steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary --no-roofline"
timeout 300 python bench.py $B > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "bench exit=$?"; summ gpurun_out/bench12.json; tail -2 gpurun_out/bench12.err
DUPL_COLSUM_2PASS=1 timeout 300 python bench.py $B > gpurun_out/bench12_2pass.json 2> gpurun_out/bench12_2pass.err; echo "bench(2pass) exit=$?"; summ gpurun_out/bench12_2pass.json
DUPL_SMEM_CARVEOUT=1 timeout 300 python bench.py $B > gpurun_out/bench12_carve.json 2> gpurun_out/bench12_carve.err; echo "bench(carveout) exit=$?"; summ gpurun_out/bench12_carve.json
timeout 300 python bench.py $B > gpurun_out/bench12_b.json 2> gpurun_out/bench12_b.err; echo "bench(repeat) exit=$?"; summ gpurun_out/bench12_b.json
timeout 200 python tools/ncu_step.py --table --gaps > gpurun_out/r02_train_step_phaseB_kernels_d.txt 2> gpurun_out/ncu_table.err; sed -n '/^GAPS/,$p' gpurun_out/r02_train_step_phaseB_kernels_d.txt | cut -c1-160; tail -3 gpurun_out/ncu_table.err
