#!/bin/bash
# N GPUs (default 8): the training step with the gradient average as one all-reduce after the backward vs chunked and
# overlapped with it.  Every run under a tight timeout.
N=${N:-8}
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 \
     bench.py --gpus $N --steps 15 --warmup 4 --no-secondary --no-cpu-baseline --no-roofline > gpurun_out/n${N}_$name.json 2> gpurun_out/n${N}_$name.err
  echo "n$N[$name] exit=$? $(grep '^{' gpurun_out/n${N}_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s e2e', round(d['e2e']['value'],1), 'sync', d['ranks_in_sync'], d['params_checked'])" 2>&1 | tail -1)"
  grep -E "Error|error" gpurun_out/n${N}_$name.err | tail -2
}
run one_chunk DUPL_NCCL_MAX_CTAS=0 DUPL_GRAD_CHUNK_ELEMS=200000000
run ctas8 DUPL_NCCL_MAX_CTAS=8
run ctas4_nores DUPL_NCCL_MAX_CTAS=4 DUPL_COMM_SMS=0
run ctas16 DUPL_NCCL_MAX_CTAS=16 DUPL_COMM_SMS=16
