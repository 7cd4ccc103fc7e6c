#!/bin/bash
# Round 2, GPU call 15: default bench with the in-step non-convolution profile (BatchNorm passes, loss, SGD, weight re-tiling) and the
# collector quiesced after the warm-up; full GPU suite on the final library.
set -x
mkdir -p gpurun_out
export PCB_REPORT_DIR=$PWD/gpurun_out
timeout 300 python bench.py --steps 50 --warmup 5 --profile-json gpurun_out/r2c15_profile.json > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err
( time timeout 1200 python -m pytest tests -m gpu -q ) 2>&1 | tail -12 > gpurun_out/r2c15_pytest.txt
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2c15_smoke.txt 2>&1
