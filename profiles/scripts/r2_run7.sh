#!/bin/bash
# Round 2, GPU call 7: conflict-free A-tile stores, batched weight tiling, golden loss-curve replay; full suite + bench.
set -x
mkdir -p gpurun_out
export PCB_REPORT_DIR=$PWD/gpurun_out
timeout 200 python profiles/microbench_split.py --levels 0,1,2 --shapes 96x96,128x128 --only fwd > gpurun_out/r2c7_microbench.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -q ) 2>&1 | tail -12 > gpurun_out/r2c7_pytest.txt
timeout 300 python bench.py --steps 50 --warmup 5 --profile-json gpurun_out/r2c7_profile.json > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err
timeout 120 ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_tcgen05_split -s 1 -c 1 --csv --log-file gpurun_out/r2c7_conflicts.csv \
  python profiles/microbench_split.py --levels 0 --shapes 96x96 --only fwd > gpurun_out/r2c7_ncu.log 2>&1
ls -la gpurun_out | tail -8
