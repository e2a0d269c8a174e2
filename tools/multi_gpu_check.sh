#!/bin/bash
# N-GPU re-check of the session-2 kernels: gradient equivalence + weak-scaling bench + the gpu-marked 2-GPU test.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 150 $TR bench.py --gpus $N --check > gpurun_out/r2h_check_n$N.json 2> gpurun_out/r2h_check_n$N.err; cat gpurun_out/r2h_check_n$N.json
timeout 200 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2h_bench_n$N.json 2> gpurun_out/r2h_bench_n$N.err; cut -c1-260 gpurun_out/r2h_bench_n$N.json; tail -1 gpurun_out/r2h_bench_n$N.err | cut -c1-200
timeout 200 python -m pytest tests -m gpu -q -k multi_gpu 2>&1 | tail -2
