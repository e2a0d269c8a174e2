#!/bin/bash
# Round-end evidence run on one B200: parity tests, the contract bench, a launch list of one training step and one
# `ncu --set full` capture per main kernel.  Outputs land in gpurun_out/ (copied into profiles/ by hand).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 500 python bench.py 2>gpurun_out/bench_err.log > gpurun_out/r1_bench_n1_v2.json
cut -c1-400 gpurun_out/r1_bench_n1_v2.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
  --log-file gpurun_out/r1_launches_train_step_v4.csv python tools/prof_target.py > gpurun_out/pf.log 2>&1
python tools/launch_summary.py gpurun_out/r1_launches_train_step_v4.csv > gpurun_out/r1_launches_train_step_v4_summary.txt 2>/dev/null
head -8 gpurun_out/r1_launches_train_step_v4_summary.txt
i=0
for spec in "convgemm_kernel<__half, 0>:30:1" "dgradT_fused_kernel:30:1" "wgrad_kernel:30:3" "head_rows_kernel:0:1"; do
  i=$((i+1))
  k=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $skip -c $cnt \
    -o gpurun_out/r1_v2_full_$i -f python tools/prof_target.py > gpurun_out/ncu_full_$i.log 2>&1
  tail -1 gpurun_out/ncu_full_$i.log
done
