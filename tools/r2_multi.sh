#!/bin/bash
# N-GPU run: gradient-equivalence check, weak-scaling bench, strong-scaling (global batch 128, micro-batched) bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 120 $TR bench.py --gpus $N --check > gpurun_out/r2_check_n$N.json 2> gpurun_out/r2_check_n$N.err; cat gpurun_out/r2_check_n$N.json
timeout 150 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; cut -c1-260 gpurun_out/r2_bench_n$N.json; tail -1 gpurun_out/r2_bench_n$N.err | cut -c1-200
timeout 200 $TR bench.py --gpus $N --steps 3 --warmup 3 --global-batch 128 > gpurun_out/r2_bench_strong128_n$N.json 2> gpurun_out/r2_bench_strong128_n$N.err; cut -c1-260 gpurun_out/r2_bench_strong128_n$N.json; tail -1 gpurun_out/r2_bench_strong128_n$N.err | cut -c1-200
