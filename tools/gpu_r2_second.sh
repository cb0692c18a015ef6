#!/bin/bash
# Round 2, second GPU call (2 GPUs): arena path on 1 GPU (train tests), the COCO tie diagnostics, NCCL test, and the
# training step at N = 2 for a few settings of the overlapped gradient average.
mkdir -p gpurun_out
export DUPL_ORACLE_CACHE=/tmp/dupl_oracle_cache
LOG=gpurun_out/tests2.log; : > $LOG
for f in tests/test_gpu_train.py tests/test_gpu_ddp_nccl.py "tests/test_gpu_baseline_sizes.py -k coco"; do
  echo "=== $f" >> $LOG; timeout 1500 python -m pytest -q --no-header -p no:cacheprovider $f -m gpu -s >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "^===|exit=|passed|failed|Error|^E |^coco|^\{" $LOG | cut -c1-900 | tail -40
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 \
     bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/n2_$name.json 2> gpurun_out/n2_$name.err
  echo "n2[$name] exit=$? $(grep '^{' gpurun_out/n2_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s sync', d['ranks_in_sync'])" 2>&1)"
  tail -2 gpurun_out/n2_$name.err
}
timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/n1_arena.json 2> gpurun_out/n1_arena.err
echo "n1 $(grep '^{' gpurun_out/n1_arena.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2),'ms')" 2>&1)"; tail -2 gpurun_out/n1_arena.err
run ctas4 DUPL_NCCL_MAX_CTAS=4
run ctas8 DUPL_NCCL_MAX_CTAS=8
run ctas4_nores DUPL_NCCL_MAX_CTAS=4 DUPL_COMM_SMS=0
run default_group DUPL_NCCL_MAX_CTAS=0
run one_chunk DUPL_NCCL_MAX_CTAS=0 DUPL_GRAD_CHUNK_ELEMS=200000000
