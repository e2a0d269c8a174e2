#!/bin/bash
# Round-2 GPU call 1: parity at depth, the contract bench, gate-gradient schedule A/B, GPU-eager bar, in-situ trace.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_box_a.txt
timeout 900 python -m pytest tests -q -m gpu --durations=12 2>&1 | tail -60 > gpurun_out/r2_tests_a.log
tail -5 gpurun_out/r2_tests_a.log
timeout 600 python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
cut -c1-300 gpurun_out/r2_bench_a.json; tail -3 gpurun_out/r2_bench_a.err
timeout 400 python tools/ab_fuse.py "CLSTM_WG_GATE=0" "CLSTM_WG_GATE=2" "CLSTM_WG_GATE=1" > gpurun_out/r2_ab_wggate.log 2>&1
cat gpurun_out/r2_ab_wggate.log | grep -v "^$" | tail -14
timeout 400 python bench.py --impl eager-gpu --steps 3 > gpurun_out/r2_eager_gpu.json 2> gpurun_out/r2_eager_gpu.err
cut -c1-400 gpurun_out/r2_eager_gpu.json; tail -2 gpurun_out/r2_eager_gpu.err
timeout 200 python tools/trace_step.py 2 > gpurun_out/r2_trace_a.txt 2>&1
head -30 gpurun_out/r2_trace_a.txt
