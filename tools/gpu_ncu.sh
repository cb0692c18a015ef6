#!/bin/bash
# ncu evidence (B200_PROFILING.md recipe): launch lists (device-time shares) + --set full captures of the dominant kernels.
# Every ncu command runs under a tight timeout: a capture costs GPU-minutes.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-step --no-graph"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 880 -c 440 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 60 -c 4 -o gpurun_out/prof_gemm -f $B > gpurun_out/ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 12 -c 2 -o gpurun_out/prof_attn -f $B > gpurun_out/ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"par_propagate|par_affinity|mscam_kernel" -s 4 -c 8 -o gpurun_out/prof_par -f $B > gpurun_out/ncu_par.log 2>&1
if [ -n "$WITH_TRAIN" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2500 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1
fi
if [ -n "$WITH_CRF" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"crf" -c 400 --csv --log-file gpurun_out/launches_crf.csv \
  python bench.py --workload crf_sweep --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_crf.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep
