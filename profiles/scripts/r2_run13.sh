#!/bin/bash
# Round 2, GPU call 13: per-tile fixed cost of the forward kernel -- PCB_TC5_DEBUG=16 skips the main loop (prologue + epilogue + CTA turnover
# only); default = the prologue with its table loads issued back to back.
set -x
mkdir -p gpurun_out
MB="python profiles/microbench_split.py --levels 0,1,2,3,4 --shapes 96x96,128x128,256x256 --only fwd"
timeout 200 $MB > gpurun_out/r2c13_mb_default.txt 2>&1
PCB_TC5_DEBUG=16 timeout 200 $MB > gpurun_out/r2c13_mb_debug16.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_c1.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2c13_pytest.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c13_bench.json 2> gpurun_out/r2c13_bench.err
