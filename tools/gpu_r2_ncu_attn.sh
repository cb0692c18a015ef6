#!/bin/bash
mkdir -p gpurun_out
python tools/attn_bench.py
TRAIN_ONLY=1 REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_dq|attn_bwd_dkv" -s 6 -c 2 -o gpurun_out/prof_attn_bwd -f python tools/attn_bench.py > gpurun_out/ncu_attn_bwd.log 2>&1
tail -3 gpurun_out/ncu_attn_bwd.log
python tools/ncu_summary.py gpurun_out/prof_attn_bwd.ncu-rep 2>/dev/null | tail -8
