#!/bin/bash
# vectorised split kernel: bit-identity with the tile kernel (train test compares MN / non-MN gradients), A/B
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_train.py -q --no-header -p no:cacheprovider -m gpu -x 2>&1 | tail -3
timeout 200 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from dupl_b200 import train
torch.manual_seed(0)
for R, Cc, gelu in [(3140, 768, False), (3140, 3072, True), (3140, 2304, False), (3136, 512, False), (34, 768, False), (130, 3072, True)]:
    src = torch.randn(R, Cc, device="cuda")
    pre = torch.randn(R, Cc, device="cuda") if gelu else None
    outs = {}
    for mode in ("1", "0"):
        os.environ["DUPL_SPLIT_ROWS"] = mode
        # the switch is read once per process: run the other mode in a subprocess instead
        break
    (hi, lo), _, cs = train.split_transpose(src, R, Cc, want_t=False, want_colsum=True, gelu_pre=pre)
    (hi2, lo2), (thi, tlo), cs2 = train.split_transpose(src, R, Cc, want_t=True, want_colsum=True, gelu_pre=pre)   # tile kernel
    torch.cuda.synchronize()
    ref = src * (0.5 * (1 + torch.erf(pre * 0.7071067811865476)) + pre * torch.exp(-0.5 * pre * pre) * 0.3989422804014327) if gelu else src
    print(R, Cc, gelu, "planes equal", torch.equal(hi, hi2) and torch.equal(lo, lo2), "colsum equal", torch.equal(cs, cs2),
          "colsum err", float((cs.double() - ref.double().sum(0)).abs().max() / ref.double().sum(0).abs().max()))
PY
summ() { grep '^{' $1 | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('$1', round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s; loss', d.get('loss'), 'clk', d.get('clocks',{}).get('sm_mhz'))
"; }
B="--steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary --no-roofline"
timeout 200 python bench.py $B > gpurun_out/bench16.json 2> gpurun_out/bench16.err; echo "exit=$?"; summ gpurun_out/bench16.json
DUPL_SPLIT_ROWS=0 timeout 200 python bench.py $B > gpurun_out/bench16_tile.json 2> gpurun_out/bench16_tile.err; echo "exit=$?"; summ gpurun_out/bench16_tile.json
timeout 200 python bench.py $B > gpurun_out/bench16_b.json 2> gpurun_out/bench16_b.err; echo "exit=$?"; summ gpurun_out/bench16_b.json
timeout 100 python tools/ncu_step.py --table 2>/dev/null | grep -E 'split_rows|split_transpose|Self CUDA time total' | cut -c1-40,150-230
