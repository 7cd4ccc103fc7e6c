#!/bin/bash
# Round 2, GPU call 5: wgrad table-prefetch ring, PDL on by default, conv variants (256-row tiles, 4-stage prefetch), microbench, ncu.
set -x
mkdir -p gpurun_out
export PCB_REPORT_DIR=$PWD/gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2c5_pytest.txt
PCB_TC5_CFG=4 timeout 600 python -m pytest tests/test_gpu_c1.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2c5_pytest_cfg4.txt
timeout 300 python profiles/microbench_split.py --levels 0,1,2,3,4 --shapes 96x96,128x128,256x256 > gpurun_out/r2c5_microbench.txt 2>&1
PCB_TC5_CFG=4 timeout 300 python profiles/microbench_split.py --levels 0,1 --shapes 96x96,128x128 --only fwd > gpurun_out/r2c5_microbench_cfg4.txt 2>&1
PCB_TC5_CFG=5 timeout 300 python profiles/microbench_split.py --levels 0,1 --shapes 96x96,128x128 --only fwd > gpurun_out/r2c5_microbench_cfg5.txt 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile-json gpurun_out/r2c5_profile.json > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
PCB_TC5_CFG=4 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c5_bench_cfg4.json 2>> gpurun_out/r2c5_bench.err
PCB_TC5_CFG=5 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c5_bench_cfg5.json 2>> gpurun_out/r2c5_bench.err
PCB_PDL=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c5_bench_nopdl.json 2>> gpurun_out/r2c5_bench.err
timeout 300 python profiles/grad_precision_ab.py 0.5 > gpurun_out/r2c5_grad_ab_05.json 2> gpurun_out/r2c5_grad_ab_05.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tcgen05 -s 1 -c 1 -o gpurun_out/r2c5_wgrad_block8 \
  python profiles/microbench_split.py --levels 0 --shapes 96x96 --only wgrad > gpurun_out/r2c5_ncu_wg0.log 2>&1
ls -la gpurun_out | tail -12
