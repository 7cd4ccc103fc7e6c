#!/bin/bash
# Round 2, GPU call 14: weight-gradient kernel with 2 offsets per CTA and two CTAs per SM (PCB_WG_GROUP=2) against 4 offsets / one CTA per SM.
set -x
mkdir -p gpurun_out
MB="python profiles/microbench_split.py --levels 0,1,2,3,4 --shapes 96x96,128x128,256x256 --only wgrad"
for g in 4 2; do PCB_WG_GROUP=$g timeout 200 $MB > gpurun_out/r2c14_mb_group$g.txt 2>&1; done
PCB_WG_GROUP=2 PCB_WG_FENCE=consumer timeout 200 $MB > gpurun_out/r2c14_mb_group2_consumer.txt 2>&1
PCB_WG_GROUP=2 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_c1.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2c14_pytest_group2.txt
for g in 4 2; do PCB_WG_GROUP=$g timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c14_bench_group$g.json 2> gpurun_out/r2c14_bench_group$g.err; done
