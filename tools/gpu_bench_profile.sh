#!/bin/bash
# bench + ncu launch list + full captures of the dominant kernels. Outputs under gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/bench.err
python bench.py --steps 10 --warmup 3 --fuse-students --no-cpu-baseline > gpurun_out/bench_fused.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 320 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 30 -c 3 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 10 -c 2 -o gpurun_out/prof_attn -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"par_propagate|par_affinity|mscam_kernel" -s 4 -c 4 -o gpurun_out/prof_par -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_par.log 2>&1
cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.json; cat gpurun_out/bench_fused.json; tail -3 gpurun_out/bench.err
