#!/bin/bash
# Round evidence run on one B200: parity tests, the contract bench, a launch list of one training step and one
# `ncu --set full` capture per main kernel.  Outputs land in gpurun_out/ (summaries are copied into profiles/).
set -u
cd "$(dirname "$0")/.."
R=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --durations=5 2>&1 | grep -v "^$" | cut -c1-2000 | tail -40 > gpurun_out/${R}_gpu_tests.log
tail -3 gpurun_out/${R}_gpu_tests.log
timeout 300 python tools/margin_case.py 1 12 24 12 64 12 256 256 2 3 > gpurun_out/${R}_margins_256px_depth36.log 2>&1; cat gpurun_out/${R}_margins_256px_depth36.log
timeout 500 python bench.py 2>gpurun_out/${R}_bench_err.log > gpurun_out/${R}_bench_n1.json
cut -c1-300 gpurun_out/${R}_bench_n1.json; tail -2 gpurun_out/${R}_bench_err.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
  --log-file gpurun_out/${R}_launches_train_step.csv python tools/prof_target.py > gpurun_out/pf.log 2>&1
python tools/launch_summary.py gpurun_out/${R}_launches_train_step.csv > gpurun_out/${R}_launches_train_step_summary.txt 2>/dev/null
head -12 gpurun_out/${R}_launches_train_step_summary.txt
i=0
for spec in "convgemm_kernel:3:1" "cellstep_pair_kernel:30:1" "dgradT_fused:30:2" "wgrad_kernel:30:2" "head_rows_kernel:0:1" "gate_grad_kernel:0:1"; do
  i=$((i+1))
  k=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $skip -c $cnt \
    -o gpurun_out/${R}_full_$i -f python tools/prof_target.py > gpurun_out/ncu_full_$i.log 2>&1
  tail -1 gpurun_out/ncu_full_$i.log
done
# the persistent small-shape chain
PROF_SMALL=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:rollout_persist" -s 1 -c 1 \
  -o gpurun_out/${R}_full_persist -f python tools/prof_target.py > gpurun_out/ncu_full_p.log 2>&1
tail -1 gpurun_out/ncu_full_p.log
# gpurun merges at most 64 MiB back: summarise on the box, keep only the fused-dgrad report (source view)
python tools/ncu_summary.py gpurun_out/${R}_full_*.ncu-rep > gpurun_out/${R}_ncu_full_summary.csv 2>/dev/null
cat gpurun_out/${R}_ncu_full_summary.csv | cut -c1-400
for f in gpurun_out/${R}_full_*.ncu-rep; do case "$f" in *_full_3.ncu-rep) ;; *) rm -f "$f";; esac; done
ls -la gpurun_out/
