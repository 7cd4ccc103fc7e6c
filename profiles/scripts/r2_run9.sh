#!/bin/bash
# Round 2, GPU call 9: proxy fence on the MMA thread (PCB_TC5_FENCE=consumer) vs on every producer: kernel microbenchmarks, parity tests, bench.
set -x
mkdir -p gpurun_out
for f in producer consumer; do
  PCB_TC5_FENCE=$f timeout 300 python profiles/microbench_split.py --levels 0,1,2,3,4 --shapes 96x96,128x128,256x256 > gpurun_out/r2c9_microbench_$f.txt 2>&1
done
PCB_TC5_FENCE=consumer timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_c1.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2c9_pytest_consumer.txt
for f in producer consumer; do
  PCB_TC5_FENCE=$f timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c9_bench_$f.json 2> gpurun_out/r2c9_bench_$f.err
done
