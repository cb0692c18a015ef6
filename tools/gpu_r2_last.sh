#!/bin/bash
# last sanity check of the final build: smoke() and a short bench line
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-secondary --no-roofline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s loss', d['loss'], d['clocks']['sm_mhz'])"
