#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 4000 --csv --log-file gpurun_out/launches_train.csv \
    python tools/bench_train.py --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/ncu_train.log
