#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_crf.py tests/test_gpu_eval_sweep.py -q --no-header -p no:cacheprovider -m gpu 2>&1 | tail -4
timeout 300 python - <<'PY'
import sys, json, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, bench
h = bench.Harness(None)
out, _ = bench.measure_crf(h, 8)
print("CRF", json.dumps(out))
PY
