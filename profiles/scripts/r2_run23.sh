#!/bin/bash
# Round 2, GPU call 23 (gpurun --gpus 8): the benchmark at N=8 over NCCL / NVSwitch (one rank per GPU, torchrun).
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2c23_bench_n8.json 2> gpurun_out/r2c23_bench_n8.err
tail -3 gpurun_out/r2c23_bench_n8.err
