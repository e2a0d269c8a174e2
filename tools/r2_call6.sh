#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^$" | cut -c1-1800 | tail -30
timeout 200 python tools/trace_small.py 2>&1 | head -3
CLSTM_GRAPH=0 timeout 200 python tools/trace_small.py 2>&1 | head -3
timeout 100 python -c "
import sys; sys.path.insert(0,'tools'); import quick_bench as q
q.run(2,4,4,32,64,False,iters=50); q.run(2,4,4,32,64,True,iters=50); q.run(1,12,24,64,128,False,iters=20); q.run(1,12,24,64,128,True,iters=20)" 2>&1 | tail -4
CLSTM_GRAPH=0 CLSTM_PERSIST=0 timeout 100 python -c "
import sys; sys.path.insert(0,'tools'); import quick_bench as q
q.run(2,4,4,32,64,False,iters=50); q.run(2,4,4,32,64,True,iters=50); q.run(1,12,24,64,128,False,iters=20); q.run(1,12,24,64,128,True,iters=20)" 2>&1 | tail -4
