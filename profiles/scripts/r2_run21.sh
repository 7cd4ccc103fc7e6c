#!/bin/bash
# Round 2, GPU call 21: pipelined iteration loop (Trainer.iter_losses) -- trainer / DDP / voxel tests, bench (e2e now through it), hardest loss.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_ddp.py tests/test_gpu_voxel.py tests/test_gpu_semseg.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2c21_pytest.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c21_bench.json 2> gpurun_out/r2c21_bench.err
timeout 300 python bench.py --loss hardest --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c21_bench_hardest.json 2>> gpurun_out/r2c21_bench.err
tail -5 gpurun_out/r2c21_bench.err
