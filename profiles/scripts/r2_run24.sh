#!/bin/bash
# Round 2, GPU call 24: final commit -- full GPU suite, default-shaped bench (per-step CUDA events from the trainer), C4 inference, smoke.
set -x
mkdir -p gpurun_out
export PCB_REPORT_DIR=$PWD/gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) 2>&1 | tail -8 > gpurun_out/r2c24_pytest.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c24_bench.json 2> gpurun_out/r2c24_bench.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/r2c24_bench_c4.json 2>> gpurun_out/r2c24_bench.err
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2c24_smoke.txt 2>&1
tail -3 gpurun_out/r2c24_pytest.txt gpurun_out/r2c24_smoke.txt
