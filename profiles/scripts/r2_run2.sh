#!/bin/bash
# Round 2, GPU call 2: unit-call executor + fused BatchNorm statistics + side-stream coordinate build + overlapped DDP all-reduce.
set -x
mkdir -p gpurun_out
export PCB_REPORT_DIR=$PWD/gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2c2_pytest.txt
timeout 300 python bench.py --steps 50 --warmup 5 --profile-json gpurun_out/r2c2_profile.json > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
PCB_SEPARATE_STATS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c2_bench_sepstats.json 2>> gpurun_out/r2c2_bench.err
PCB_COORDS_STREAM=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c2_bench_nostream.json 2>> gpurun_out/r2c2_bench.err
timeout 300 python bench.py --workload c0 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c2_bench_c0.json 2>> gpurun_out/r2c2_bench.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/r2c2_bench_c4.json 2>> gpurun_out/r2c2_bench.err
timeout 300 python bench.py --loss hardest --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c2_bench_hardest.json 2>> gpurun_out/r2c2_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c2_bench_reference.json 2>> gpurun_out/r2c2_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2200 -c 900 --csv \
  --log-file gpurun_out/r2c2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_ncu_bench.log 2>&1
ls -la gpurun_out | tail -15
