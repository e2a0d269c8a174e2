#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --durations=8 2>&1 | grep -v "^$" | cut -c1-2500 > gpurun_out/r2_tests_b.log
tail -60 gpurun_out/r2_tests_b.log
timeout 300 python tools/quick_bench.py > gpurun_out/r2_quick_b.log 2>&1; cat gpurun_out/r2_quick_b.log
CLSTM_PERSIST=0 timeout 100 python -c "
import sys; sys.path.insert(0,'tools'); import quick_bench as q
q.run(2,4,4,32,64,False,iters=50); q.run(2,4,4,32,64,True,iters=50)" 2>&1 | tail -3
timeout 100 python -c "
import sys; sys.path.insert(0,'tools'); import quick_bench as q
q.run(2,4,4,32,64,False,iters=50); q.run(2,4,4,32,64,True,iters=50); q.run(1,12,24,64,128,False,iters=20)" 2>&1 | tail -3
CLSTM_PERSIST=0 timeout 100 python -c "
import sys; sys.path.insert(0,'tools'); import quick_bench as q
q.run(1,12,24,64,128,False,iters=20)" 2>&1 | tail -3
