#!/bin/bash
# Round checkpoint on the B200 box: full GPU test suite (one process per file), bench.py with breakdown,
# training-step bench with a torch.profiler kernel table, and ncu launch lists of both.  Outputs: gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/tests.log; : > $LOG
for f in tests/test_gpu_*.py; do
  echo "=== $f" >> $LOG; timeout ${TEST_TIMEOUT:-300} python -m pytest -q -x --no-header -p no:cacheprovider $f -m gpu >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|assert" $LOG | tail -40
timeout 400 python bench.py --steps 10 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 400 python tools/bench_train.py --steps 5 --warmup 3 --profile > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train exit=$?"
cat gpurun_out/bench_train.json
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 4000 --csv --log-file gpurun_out/launches_train.csv \
    python tools/bench_train.py --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1
fi
