#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/tests.log; : > $LOG
for f in tests/test_gpu_cam_par.py tests/test_gpu_golden.py; do
  echo "=== $f" >> $LOG; timeout 400 python -m pytest -q -x --no-header -p no:cacheprovider $f -m gpu >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|assert|^E " $LOG | tail -40
timeout 400 python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline --no-train-step > gpurun_out/bench_cam_par.json 2> gpurun_out/bench_cam_par.err; echo "bench exit=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_cam_par.json'))
print('img/s', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'gemm TF', round(d['roofline']['achieved'],1), d['breakdown_ms'])"
tail -3 gpurun_out/bench_cam_par.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"par_propagate|mscam_kernel" -s 4 -c 6 -o gpurun_out/prof_par3 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-step --no-graph > gpurun_out/ncu_par3.log 2>&1
tail -1 gpurun_out/ncu_par3.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"crf" -c 400 --csv --log-file gpurun_out/launches_crf.csv \
  python bench.py --workload crf_sweep --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_crf.log 2>&1
tail -1 gpurun_out/ncu_crf.log
