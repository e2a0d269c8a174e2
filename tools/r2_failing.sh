#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "unrolled or bf16 or weights_x8 or stale_graph or cloudgan" 2>&1 | grep -v "^$" | cut -c1-1500 > gpurun_out/r2_failing.log
tail -150 gpurun_out/r2_failing.log
timeout 300 python bench.py --impl eager-gpu --steps 3 > gpurun_out/r2_eager_gpu.json 2> gpurun_out/r2_eager_gpu.err
cut -c1-600 gpurun_out/r2_eager_gpu.json; tail -2 gpurun_out/r2_eager_gpu.err
