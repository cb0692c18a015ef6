#!/bin/bash
# Final numbers, 1 GPU: the driver's two bench commands, phase C, and the ncu evidence of the final kernels
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/bench_default_final.err; echo "default bench exit=$?"
timeout 400 python bench.py --impl reference > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/bench_reference_final.err; echo "reference bench exit=$?"
timeout 300 python bench.py --phase C --steps 15 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-secondary --no-roofline > gpurun_out/r02_bench_phaseC_final.json 2> gpurun_out/bench_phaseC_final.err; echo "phase C exit=$?"
python - <<'PY'
import json
for f in ("r02_bench_default_final", "r02_bench_reference_final", "r02_bench_phaseC_final"):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
    except Exception as e:
        print(f, "NO LINE", e); continue
    print(f, round(d.get("ms_per_step", 0), 2), "ms", round(d.get("value", 0), 3), d.get("unit"), "e2e", d.get("e2e", {}).get("value"), "clocks", d.get("clocks"))
    for k in ("roofline", "cpu_baseline", "reference_gpu", "cam_par", "crf"):
        if k in d:
            print("   ", k, json.dumps(d[k])[:330])
PY
tail -2 gpurun_out/bench_default_final.err
NCU="ncu --profile-from-start off --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum --graph-profiling node --csv --log-file gpurun_out/r02_launches_train_step_final.csv python tools/ncu_step.py > gpurun_out/ncu_list.log 2>&1; echo "list exit=$?"; wc -l gpurun_out/r02_launches_train_step_final.csv
cap() {  # name, regex, skip, count
  timeout 240 $NCU --set full --import-source on --graph-profiling node -k regex:"$2" -s $3 -c $4 -o gpurun_out/r02_prof_$1 -f python tools/ncu_step.py > gpurun_out/ncu_$1.log 2>&1
  echo "cap[$1] exit=$?"; tail -1 gpurun_out/ncu_$1.log | cut -c1-160
}
cap hbm2 "mscam_cluster|split_transpose|par_propagate|par_affinity|adamw_update|cam_contract|layernorm_bwd_kernel" 0 12
cap gemm2 "gemm_bf16x3" 100 8
cap attnbwd "attn_bwd_dq|attn_bwd_dkv" 4 2
for f in hbm2 gemm2 attnbwd; do python tools/ncu_summary.py gpurun_out/r02_prof_$f.ncu-rep > gpurun_out/r02_ncu_$f.md 2>/dev/null; cut -c1-230 gpurun_out/r02_ncu_$f.md; done
timeout 200 python tools/ncu_step.py --table > gpurun_out/r02_train_step_phaseB_kernels_final.txt 2> gpurun_out/ncu_table.err; tail -3 gpurun_out/r02_train_step_phaseB_kernels_final.txt
