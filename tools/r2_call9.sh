#!/bin/bash
cd "$(dirname "$0")/.."
for w in 2 4; do
CLSTM_FUSE_WORKERS=$w timeout 300 python -m pytest tests -q -m gpu -x -k "benched_depth_12 or stress_weights_x3 or full_width or linear_in_loss or rollout_golden or three_layers" 2>&1 | tail -2
done
timeout 400 python tools/ab_fuse.py "CLSTM_FUSE_WORKERS=0" "CLSTM_FUSE_WORKERS=2" "CLSTM_FUSE_WORKERS=4" 2>&1 | grep -v "^$"
for v in 0 2 4; do CLSTM_FUSE_WORKERS=$v timeout 120 python tools/kernel_bench.py 16 dgrad_fused 2>&1 | tail -1; done
