#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 320 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 10 -c 1 -o gpurun_out/prof_attn -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
