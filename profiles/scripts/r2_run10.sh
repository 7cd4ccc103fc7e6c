#!/bin/bash
# Round 2, GPU call 10 (gpurun --gpus 2): the data-parallel path on real NCCL -- bench at N=2 (PointInfoNCE and hardest-contrastive), the
# reference arm under torchrun (rank 0 works, rank 1 exits), the DDP tests with two devices visible, and N=1 on the same box for the ratio.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2c10_bench_n2.json 2> gpurun_out/r2c10_bench_n2.err
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c10_bench_n1.json 2> gpurun_out/r2c10_bench_n1.err
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --loss hardest --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c10_bench_n2_hardest.json 2> gpurun_out/r2c10_bench_n2_hardest.err
timeout 600 $TR --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2c10_bench_n2_reference.json 2> gpurun_out/r2c10_bench_n2_reference.err
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2c10_pytest_ddp.txt
tail -3 gpurun_out/r2c10_*.err
