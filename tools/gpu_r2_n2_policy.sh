#!/bin/bash
# gradient-average policy at 2 GPUs after the round's backward changes: one all-reduce vs chunked + overlapped (NCCL CTAs / reserved SMs)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
B="bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary --no-roofline"
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 200 $TR $B > gpurun_out/n2p_$tag.json 2> gpurun_out/n2p_$tag.err; rc=$?
  python - "$tag" "$rc" <<'PY'
import json, sys
tag, rc = sys.argv[1], sys.argv[2]
try:
    d = [json.loads(l) for l in open(f"gpurun_out/n2p_{tag}.json") if l.startswith("{")][-1]
    print(tag, "rc", rc, round(d["ms_per_step"], 2), "ms", round(d["value"], 1), "img/s sync", d.get("ranks_in_sync"), "clk", d.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print(tag, "rc", rc, "NO LINE", e)
PY
}
run one_chunk_a DUPL_GRAD_OVERLAP=0
run overlap16_a DUPL_GRAD_OVERLAP=1
run overlap8 DUPL_GRAD_OVERLAP=1 DUPL_NCCL_MAX_CTAS=8 DUPL_COMM_SMS=8
run overlap24 DUPL_GRAD_OVERLAP=1 DUPL_NCCL_MAX_CTAS=24 DUPL_COMM_SMS=24
run one_chunk_b DUPL_GRAD_OVERLAP=0
run overlap16_b DUPL_GRAD_OVERLAP=1
