#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/tests.log; : > $LOG
for f in ${TEST_FILES:-tests/test_gpu_cam_par.py tests/test_gpu_golden.py tests/test_gpu_train.py}; do
  echo "=== $f" >> $LOG; timeout ${TEST_TIMEOUT:-400} python -m pytest -q -x --no-header -p no:cacheprovider $f -m gpu >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|assert|^E " $LOG | tail -40
timeout 400 python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/bench_cam_par.json 2> gpurun_out/bench_cam_par.err; echo "bench exit=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_cam_par.json'))
print('img/s', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'gemm TF', round(d['roofline']['achieved'],1), d['breakdown_ms'], 'train', d.get('train_step',{}).get('ms_per_step'))"
tail -3 gpurun_out/bench_cam_par.err
if [ -n "$NCU" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"par_propagate|mscam_kernel" -s 4 -c 6 -o gpurun_out/prof_par2 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-step --no-graph > gpurun_out/ncu_par2.log 2>&1
tail -2 gpurun_out/ncu_par2.log
fi
