#!/bin/bash
# Round 2, GPU call 22: what the driver runs at round end, on the final commit -- GPU suite, smoke(), default bench, reference arm.
set -x
mkdir -p gpurun_out
export PCB_REPORT_DIR=$PWD/gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8 > gpurun_out/r2c22_pytest.txt
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2c22_smoke.txt 2>&1
timeout 400 python bench.py > gpurun_out/r2c22_bench.json 2> gpurun_out/r2c22_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c22_bench_reference.json 2>> gpurun_out/r2c22_bench.err
tail -3 gpurun_out/r2c22_pytest.txt gpurun_out/r2c22_smoke.txt
