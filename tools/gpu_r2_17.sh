#!/bin/bash
# LayerNorm backward with 8 rows per block: gradients vs the oracle + A/B
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_train.py -q --no-header -p no:cacheprovider -m gpu -x -k "student_forward or phase_b or in_place or captured" 2>&1 | tail -2
summ() { grep '^{' $1 | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('$1', round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s; loss', d.get('loss'), 'clk', d.get('clocks',{}).get('sm_mhz'))
"; }
B="--steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary --no-roofline"
timeout 150 python bench.py $B > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "exit=$?"; summ gpurun_out/bench17.json
DUPL_LNB_ROWS=16 timeout 150 python bench.py $B > gpurun_out/bench17_16.json 2> gpurun_out/bench17_16.err; echo "exit=$?"; summ gpurun_out/bench17_16.json
timeout 150 python bench.py $B > gpurun_out/bench17_b.json 2> gpurun_out/bench17_b.err; echo "exit=$?"; summ gpurun_out/bench17_b.json
