#!/bin/bash
# Round 2, GPU call 12: cp.async producers (PCB_TC5_CFG=7: 2 CTAs/SM x 3 slots, 8: 3 CTAs/SM x 2 slots) against the register-staged default.
set -x
mkdir -p gpurun_out
for c in 2 7 8; do PCB_TC5_CFG=$c timeout 200 python profiles/microbench_split.py --levels 0,1,2,3,4 --shapes 96x96,128x128,256x256 --only fwd > gpurun_out/r2c12_mb_cfg$c.txt 2>&1; done
PCB_TC5_CFG=7 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_c1.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2c12_pytest_cfg7.txt
PCB_TC5_CFG=8 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_c1.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2c12_pytest_cfg8.txt
for c in 2 7 8; do PCB_TC5_CFG=$c timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c12_bench_cfg$c.json 2> gpurun_out/r2c12_bench_cfg$c.err; done
