#!/bin/bash
cd "$(dirname "$0")/.."
timeout 200 python tools/trace_small.py 2>&1 | head -45
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:convgemm_kernel" -s 30 -c 1 -o gpurun_out/r2_full_1 -f python tools/prof_target.py > gpurun_out/ncu_full_1.log 2>&1
tail -1 gpurun_out/ncu_full_1.log
