#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:dgradT_fused" -s 30 -c 1 -o gpurun_out/r2_fused_sets4 -f python tools/prof_target.py > gpurun_out/ncu_fused4.log 2>&1
tail -1 gpurun_out/ncu_fused4.log
