#!/bin/bash
# Final validation, 1 GPU: the whole -m gpu suite (what the driver runs at round end) + smoke()
mkdir -p gpurun_out
export DUPL_ORACLE_CACHE=/tmp/dupl_oracle_cache
timeout 1300 python -m pytest tests -q --no-header -p no:cacheprovider -m gpu > gpurun_out/tests_final.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/tests_final.log | cut -c1-300 | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
