#!/bin/bash
# Round 2, GPU call 25 (gpurun --gpus 2): the final bench.py under torchrun at N=2 (timed loop = the trainer's loop).
set -x
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c25_bench_n2.json 2> gpurun_out/r2c25_bench_n2.err
tail -2 gpurun_out/r2c25_bench_n2.err
