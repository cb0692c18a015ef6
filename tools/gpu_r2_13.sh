#!/bin/bash
# bench line after the config / notes split; cam_contract tokens-per-warp A/B
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-secondary > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench exit=$?"
grep '^{' gpurun_out/bench13.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), d['config'], d['notes'], d['roofline']['frac'])"
tail -2 gpurun_out/bench13.err
for t in 1 2 4; do DUPL_CAM_TOK=$t timeout 200 python tools/ncu_step.py --table 2>/dev/null | grep -E 'cam_contract|Self CUDA time total' | cut -c1-40,150-230; done
timeout 200 python -m pytest tests/test_gpu_dense.py -q --no-header -p no:cacheprovider -m gpu -k "cam or encoder or mscam" 2>&1 | tail -2
DUPL_CAM_TOK=2 timeout 200 python -m pytest tests/test_gpu_dense.py tests/test_gpu_fullsize.py -q --no-header -p no:cacheprovider -m gpu -k "cam or encoder or mscam or flip" 2>&1 | tail -2
