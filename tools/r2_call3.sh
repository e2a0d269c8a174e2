#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "schedules or cfg1 or benched_depth_12 or ragged" 2>&1 | tail -5
for v in 1 2 1 2; do
CLSTM_STAGED=$v timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); import quick_bench as q
q.run(16,12,24,64,256,False,iters=5); q.run(16,12,24,64,256,True,iters=3)" 2>&1 | sed "s/^/STAGED=$v /"
done
