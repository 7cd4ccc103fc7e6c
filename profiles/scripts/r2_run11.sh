#!/bin/bash
# Round 2, GPU call 11: what bounds an iteration of the forward kernel -- timing ablations (PCB_TC5_DEBUG: 1 no gathers, 2 no weight tiles,
# 4 no MMAs; results are wrong by construction), the dedicated weight-loader warp (cfg 6) and the 256-row tiles (cfg 4) on top of the
# consumer-side fence; then the parity tests and the bench under the best candidate.
set -x
mkdir -p gpurun_out
MB="python profiles/microbench_split.py --levels 0,2,4 --shapes 96x96,256x256 --only fwd"
timeout 200 $MB > gpurun_out/r2c11_mb_default.txt 2>&1
for d in 1 2 4 3; do PCB_TC5_DEBUG=$d timeout 200 $MB > gpurun_out/r2c11_mb_debug$d.txt 2>&1; done
for c in 4 6; do PCB_TC5_CFG=$c timeout 200 python profiles/microbench_split.py --levels 0,1,2,3,4 --shapes 96x96,128x128,256x256 --only fwd > gpurun_out/r2c11_mb_cfg$c.txt 2>&1; done
PCB_TC5_CFG=6 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_c1.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2c11_pytest_cfg6.txt
for c in 2 4 6; do PCB_TC5_CFG=$c timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c11_bench_cfg$c.json 2> gpurun_out/r2c11_bench_cfg$c.err; done
