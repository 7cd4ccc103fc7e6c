#!/bin/bash
# Round 2, GPU call 26: C0 and hardest-contrastive numbers of the final commit.
set -x
mkdir -p gpurun_out
timeout 100 python bench.py --workload c0 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c26_bench_c0.json 2> gpurun_out/r2c26_bench.err
timeout 100 python bench.py --loss hardest --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c26_bench_hardest.json 2>> gpurun_out/r2c26_bench.err
