#!/bin/bash
# validate: programmatic dependent launch of the hot kernels (A/B), 16-CTA MS-CAM clusters, GMM tests at the 1e-4 bar
mkdir -p gpurun_out
export DUPL_ORACLE_CACHE=/tmp/dupl_oracle_cache
LOG=gpurun_out/tests11.log; : > $LOG
for f in tests/test_gpu_dense.py "tests/test_gpu_cam_par.py -k mscam" tests/test_gpu_gmm.py tests/test_gpu_train.py tests/test_gpu_losses.py; do
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
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "bench exit=$?"; summ gpurun_out/bench11.json; tail -3 gpurun_out/bench11.err
DUPL_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary --no-roofline > gpurun_out/bench11_nopdl.json 2> gpurun_out/bench11_nopdl.err; echo "bench(no pdl) exit=$?"; summ gpurun_out/bench11_nopdl.json; tail -2 gpurun_out/bench11_nopdl.err
DUPL_MSCAM_CLUSTER=8 timeout 300 python bench.py --steps 10 --warmup 3 --workload cam_par --no-cpu-baseline > gpurun_out/bench11_campar_cl8.json 2> gpurun_out/bench11_campar_cl8.err; grep -o '"mscam_post": {[^}]*}' gpurun_out/bench11_campar_cl8.json | head -2
timeout 200 python tools/ncu_step.py --table > gpurun_out/r02_train_step_phaseB_kernels_c.txt 2> gpurun_out/ncu_table.err; head -16 gpurun_out/r02_train_step_phaseB_kernels_c.txt | cut -c1-200 | awk '{print $1, $(NF-4), $(NF-2), $(NF-1), $NF}' | head -45
