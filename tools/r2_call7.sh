#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests -q -m gpu -x -k "schedules or benched_depth_12 or stress_weights_x3 or full_width or linear_in_loss" 2>&1 | tail -4
timeout 400 python tools/ab_fuse.py "CLSTM_FUSE_SETS=2" "CLSTM_FUSE_SETS=4" 2>&1 | grep -v "^$"
for v in 2 4; do CLSTM_FUSE_SETS=$v timeout 120 python tools/kernel_bench.py 16 dgrad_fused,dgrad 2>&1 | tail -2; done
