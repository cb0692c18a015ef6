#!/bin/bash
# validate: MN-major GEMM operands, 4-token cam_contract, single-launch MS-CAM, batched CRF splat; A/B the step time
mkdir -p gpurun_out
export DUPL_ORACLE_CACHE=/tmp/dupl_oracle_cache
LOG=gpurun_out/tests10.log; : > $LOG
for f in "tests/test_gpu_dense.py -k gemm" tests/test_gpu_cam_par.py tests/test_gpu_crf.py tests/test_gpu_train.py tests/test_gpu_dense.py tests/test_gpu_eval_sweep.py; do
echo "=== $f" >> $LOG; timeout 400 python -m pytest -q --no-header -p no:cacheprovider $f -m gpu >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|^E |^FAILED" $LOG | cut -c1-300 | tail -40
summ() { grep '^{' $1 | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('$1', round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s; e2e', round(d['e2e']['value'],1), 'loss', d.get('loss'), 'launches', d.get('gpu_launches'))
for k in ('cam_par','crf'):
    if k in d: print(' ', k, json.dumps(d[k])[:400])
hk=d.get('hbm_kernels')
if hk: print('  hbm', json.dumps({k:{a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!='model'} for k,v in hk.items()}))
"; }
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "bench exit=$?"; summ gpurun_out/bench10.json; tail -3 gpurun_out/bench10.err
DUPL_MN_MAJOR=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary --no-roofline > gpurun_out/bench10_nomn.json 2> gpurun_out/bench10_nomn.err; echo "bench(no mn) exit=$?"; summ gpurun_out/bench10_nomn.json
timeout 200 python tools/ncu_step.py --table > gpurun_out/r02_train_step_phaseB_kernels_b.txt 2> gpurun_out/ncu_table.err; head -40 gpurun_out/r02_train_step_phaseB_kernels_b.txt | cut -c1-200 | awk '{print $1, $(NF-4), $(NF-2), $(NF-1), $NF}' | head -45
