#!/bin/bash
# Round 2, GPU call 20: end-to-end throughput with the coordinate side stream at default vs high priority (same box, alternating).
set -x
mkdir -p gpurun_out
for i in 1 2; do for p in 0 -1; do
  PCB_COORDS_PRIORITY=$p timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c20_bench_prio${p}_$i.json 2> gpurun_out/r2c20_bench.err
done; done
