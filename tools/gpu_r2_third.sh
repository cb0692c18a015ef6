#!/bin/bash
# 1 GPU: pipelined attention backward (parity at all N + train tests), training GEMM shapes (grouping), bench.
mkdir -p gpurun_out
export DUPL_ORACLE_CACHE=/tmp/dupl_oracle_cache
LOG=gpurun_out/tests3.log; : > $LOG
for f in "tests/test_gpu_baseline_sizes.py -k attention" tests/test_gpu_train.py "tests/test_gpu_baseline_sizes.py -k phase_b"; do
  echo "=== $f" >> $LOG; timeout 900 python -m pytest -q --no-header -p no:cacheprovider $f -m gpu -s >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|^E |^attention|^train448" $LOG | cut -c1-700 | tail -30
timeout 300 python tools/gemm_shapes.py train 2>&1 | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench exit=$?"
grep '^{' gpurun_out/bench3.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print(round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s; e2e', round(d['e2e']['value'],1), 'roofline', round(d['roofline']['frac'],3), json.dumps(d['hbm_kernels']['attention_bwd']), json.dumps(d['hbm_kernels']['attention_fwd']))"
tail -3 gpurun_out/bench3.err
