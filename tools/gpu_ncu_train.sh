mkdir -p gpurun_out
DUPL_TRAIN_CAPTURE=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_dkv|attn_bwd_dq|gemm_bf16x3_kernel<192|gemm_bf16x3_kernel<128|split_transpose|layernorm_bwd_kernel" -s 700 -c 14 -o gpurun_out/prof_train -f python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_full.log 2>&1
tail -2 gpurun_out/ncu_train_full.log
