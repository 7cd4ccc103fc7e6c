#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_trainer.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2c19_pytest.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c19_bench.json 2> gpurun_out/r2c19_bench.err
