#!/bin/bash
mkdir -p gpurun_out
export DUPL_TEST_TRACE=1 DUPL_TEST_WATCHDOG=100 DUPL_TEST_SIZE=64
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29655 tests/nccl_worker.py > gpurun_out/nccl_worker.out 2> gpurun_out/nccl_worker.err
echo "worker exit=$?"
grep -E "RESULT|\[rank" gpurun_out/nccl_worker.out gpurun_out/nccl_worker.err | cut -c1-400 | tail -20
grep -E "File \"/|line [0-9]+ in|Thread|Timeout" gpurun_out/nccl_worker.err | head -50 | cut -c1-200
