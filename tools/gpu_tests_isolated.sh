#!/bin/bash
# Runs every GPU test file/function in its own process under a timeout so that one hung kernel cannot
# hide the results of the others.  Output: gpurun_out/tests_isolated.log
mkdir -p gpurun_out
LOG=gpurun_out/tests_isolated.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $LOG 2>&1
run() { echo "=== $*" >> $LOG; timeout 300 python -m pytest -q -x --no-header -p no:cacheprovider "$@" >> $LOG 2>&1; echo "exit=$?" >> $LOG; }
run tests/test_gpu_cam_par.py -m gpu
for t in test_split_bf16_reconstructs_to_2e_minus_16 test_gemm_f32_bias test_gemm_epilogues test_gemm_rejects_bad_arguments \
         test_layernorm_split test_attention_matches_fp64_softmax_attention test_cam_only_matches_oracle test_multi_scale_cam_matches_oracle; do
  run tests/test_gpu_dense.py -m gpu -k $t
done
grep -E "^===|exit=|passed|failed|Error|error" $LOG | tail -60
