#!/bin/bash
# Last validation of the round on 1 GPU: whole -m gpu suite, smoke(), the driver's default bench command + launch list
mkdir -p gpurun_out
export DUPL_ORACLE_CACHE=/tmp/dupl_oracle_cache
timeout 1000 python -m pytest tests -q --no-header -p no:cacheprovider -m gpu > gpurun_out/tests_final2.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/tests_final2.log | cut -c1-300 | tail -12
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 500 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/bench_default_final.err; echo "default bench exit=$?"
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02_bench_default_final.json") if l.startswith("{")][-1]
print(round(d["ms_per_step"], 2), "ms", round(d["value"], 2), d["unit"], "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"], d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "issued_tflops", "frac_issued")})
for k in ("cpu_baseline", "reference_gpu", "cam_par", "crf"):
    print(k, json.dumps(d.get(k))[:260])
print("hbm", json.dumps({k: {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != "model"} for k, v in d["hbm_kernels"].items()}))
PY
timeout 200 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --graph-profiling node --csv --log-file gpurun_out/r02_launches_train_step_final.csv python tools/ncu_step.py > gpurun_out/ncu_list.log 2>&1; echo "list exit=$?"; wc -l gpurun_out/r02_launches_train_step_final.csv
