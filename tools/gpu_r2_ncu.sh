#!/bin/bash
# Round-2 ncu evidence (B200_PROFILING.md recipe).  Every ncu command under a tight timeout.
mkdir -p gpurun_out
timeout 200 python tools/ncu_step.py --table > gpurun_out/r02_train_step_phaseB_kernels.txt 2> gpurun_out/ncu_table.err; echo "table exit=$?"
head -40 gpurun_out/r02_train_step_phaseB_kernels.txt | cut -c1-150
NCU="ncu --profile-from-start off --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum --graph-profiling node --csv --log-file gpurun_out/r02_launches_train_step.csv python tools/ncu_step.py > gpurun_out/ncu_list.log 2>&1; echo "list exit=$?"
wc -l gpurun_out/r02_launches_train_step.csv
cap() {  # name, regex, skip, count
  timeout 300 $NCU --set full --import-source on --graph-profiling node -k regex:"$2" -s $3 -c $4 -o gpurun_out/r02_prof_$1 -f python tools/ncu_step.py > gpurun_out/ncu_$1.log 2>&1
  echo "cap[$1] exit=$?"; tail -1 gpurun_out/ncu_$1.log | cut -c1-200
}
cap gemm "gemm_bf16x3" 40 6
cap attn "attention_fwd|attn_bwd" 10 4
cap hbm "par_propagate|par_affinity|mscam_kernel|layernorm|adamw_update|seg_up|split_transpose" 2 14
timeout 300 ncu --clock-control none --set full -k regex:"crf_splat|crf_blur|crf_slice" -s 30 -c 6 -o gpurun_out/r02_prof_crf -f python bench.py --workload crf_sweep --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_crf.log 2>&1; echo "crf exit=$?"
for f in gemm attn hbm crf; do python tools/ncu_summary.py gpurun_out/r02_prof_$f.ncu-rep > gpurun_out/r02_ncu_$f.md 2>/dev/null; cat gpurun_out/r02_ncu_$f.md | cut -c1-260; done
ls -la gpurun_out/*.ncu-rep | tail -6
