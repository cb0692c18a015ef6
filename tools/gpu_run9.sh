#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/tests.log; : > $LOG
echo "=== attention tests" >> $LOG; timeout 300 python -m pytest -q -x --no-header -p no:cacheprovider tests/test_gpu_dense.py -m gpu -k "attention or cam_only or multi_scale or val_forward" >> $LOG 2>&1; echo "exit=$?" >> $LOG
grep -E "^===|exit=|passed|failed|Error|assert" $LOG | tail
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json
d=json.load(open('gpurun_out/bench.json'))
print('img/s', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'gemm TF', round(d['roofline']['achieved'],1), d['breakdown_ms'])"
tail -3 gpurun_out/bench.err
timeout 900 python tools/bench_crf.py > gpurun_out/bench_crf.jsonl 2> gpurun_out/bench_crf.err; cat gpurun_out/bench_crf.jsonl; tail -3 gpurun_out/bench_crf.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 10 -c 1 -o gpurun_out/prof_attn -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
