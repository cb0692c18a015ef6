#!/bin/bash
# 8 epilogue warps in the GEMM: correctness + A/B
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_dense.py -q --no-header -p no:cacheprovider -m gpu -x 2>&1 | tail -3
timeout 200 python -m pytest tests/test_gpu_train.py -q --no-header -p no:cacheprovider -m gpu -x 2>&1 | tail -3
summ() { grep '^{' $1 | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('$1', round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s; loss', d.get('loss'), 'clk', d.get('clocks',{}).get('sm_mhz'), 'frac_issued', (d.get('roofline') or {}).get('frac_issued'))
"; }
B="--steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary"
timeout 200 python bench.py $B > gpurun_out/bench15_epi8.json 2> gpurun_out/bench15_epi8.err; echo "exit=$?"; summ gpurun_out/bench15_epi8.json
DUPL_GEMM_EPI_WARPS=4 timeout 200 python bench.py $B > gpurun_out/bench15_epi4.json 2> gpurun_out/bench15_epi4.err; echo "exit=$?"; summ gpurun_out/bench15_epi4.json
timeout 200 python bench.py $B --no-roofline > gpurun_out/bench15_epi8_b.json 2> gpurun_out/bench15_epi8_b.err; echo "exit=$?"; summ gpurun_out/bench15_epi8_b.json
DUPL_GEMM_EPI_WARPS=4 timeout 200 python bench.py $B --no-roofline > gpurun_out/bench15_epi4_b.json 2> gpurun_out/bench15_epi4_b.err; echo "exit=$?"; summ gpurun_out/bench15_epi4_b.json
