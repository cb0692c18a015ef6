#!/bin/bash
# Final validation on 2 GPUs: real-NCCL arena test, the unchanged script under torchrun x2 (M1), the data-parallel step, the sharded eval sweep
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 240 python -m pytest tests/test_gpu_ddp_nccl.py -q --no-header -p no:cacheprovider -m gpu -s > gpurun_out/tests_2gpu_nccl.log 2>&1; echo "nccl test exit=$?"; grep -E "passed|failed|skipped|^E  |RESULT|grad_mean" gpurun_out/tests_2gpu_nccl.log | cut -c1-400 | tail -6
timeout 480 python -m pytest tests/test_gpu_m1_script.py -k "2" -q --no-header -p no:cacheprovider -m gpu > gpurun_out/tests_2gpu_m1.log 2>&1; echo "m1 x2 exit=$?"; grep -E "passed|failed|skipped|^E  " gpurun_out/tests_2gpu_m1.log | cut -c1-400 | tail -6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701"
timeout 240 $TR bench.py --gpus 2 --steps 15 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary --no-roofline > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/bench_n2_final.err; echo "bench n2 exit=$?"
DUPL_GRAD_OVERLAP=1 timeout 240 $TR bench.py --gpus 2 --steps 15 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary --no-roofline > gpurun_out/r02_bench_n2_overlap_final.json 2> gpurun_out/bench_n2_overlap_final.err; echo "bench n2 (chunked + overlapped path, the N >= 4 policy) exit=$?"
timeout 300 $TR bench.py --gpus 2 --workload crf_sweep --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_crf_sweep_n2.json 2> gpurun_out/bench_crf_sweep_n2.err; echo "crf_sweep n2 exit=$?"
python - <<'PY'
import json
for f in ("r02_bench_n2_final", "r02_bench_n2_overlap_final", "r02_bench_crf_sweep_n2"):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
        print(f, d.get("metric"), round(d.get("ms_per_step", 0), 2), "ms", round(d.get("value", 0), 3), d.get("unit"), "n", d.get("n_gpus"), "e2e", d.get("e2e", {}).get("value"), "sync", d.get("ranks_in_sync"))
    except Exception as e:
        print(f, "NO LINE", e)
PY
tail -3 gpurun_out/bench_crf_sweep_n2.err
