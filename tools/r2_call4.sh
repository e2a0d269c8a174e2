#!/bin/bash
cd "$(dirname "$0")/.."
for cfg in "1 3" "2 3" "2 4" "1 2" "2 2" "1 3" "2 4"; do
set -- $cfg
CLSTM_STAGED=$1 CLSTM_STAGES=$2 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); import quick_bench as q
q.run(16,12,24,64,256,False,iters=5)" 2>&1 | sed "s/^/STAGED=$1 STAGES=$2 /"
done
