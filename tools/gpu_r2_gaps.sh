#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/ncu_step.py --table --gaps > gpurun_out/r02_gaps.txt 2> gpurun_out/gaps.err; sed -n '/^idle periods/,$p' gpurun_out/r02_gaps.txt | cut -c1-200 | head -120; tail -3 gpurun_out/gaps.err
