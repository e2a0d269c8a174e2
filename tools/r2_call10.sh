#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests -q -m gpu -x -k "schedules" 2>&1 | tail -3
for h in "50 2" "50 1"; do set -- $h
CLSTM_HYBRID=$1 CLSTM_HYBRID_WG=$2 timeout 300 python -m pytest tests -q -m gpu -x -k "benched_config_256 or full_width or linear_in_loss" 2>&1 | tail -2
done
timeout 600 python tools/ab_fuse.py "CLSTM_HYBRID=0" "CLSTM_HYBRID=50 CLSTM_HYBRID_WG=2" "CLSTM_HYBRID=50 CLSTM_HYBRID_WG=1" "CLSTM_HYBRID=35 CLSTM_HYBRID_WG=2" "CLSTM_HYBRID=65 CLSTM_HYBRID_WG=2" 2>&1 | grep -v "^$"
