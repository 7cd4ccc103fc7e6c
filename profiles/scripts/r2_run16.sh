#!/bin/bash
# Round 2, GPU call 16: final library -- GPU suite, default bench with the in-step profile, ncu launch list (time + DRAM bytes) of one step,
# ncu --set full of the forward and the weight-gradient kernel on the block8 shape.
set -x
mkdir -p gpurun_out
export PCB_REPORT_DIR=$PWD/gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) 2>&1 | tail -12 > gpurun_out/r2c16_pytest.txt
timeout 300 python bench.py --steps 50 --warmup 5 --profile-json gpurun_out/r2c16_profile.json > gpurun_out/r2c16_bench.json 2> gpurun_out/r2c16_bench.err
timeout 300 python bench.py --loss hardest --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c16_bench_hardest.json 2>> gpurun_out/r2c16_bench.err
timeout 300 python bench.py --workload c0 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c16_bench_c0.json 2>> gpurun_out/r2c16_bench.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/r2c16_bench_c4.json 2>> gpurun_out/r2c16_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2500 -c 1700 --csv --log-file gpurun_out/r2c16_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c16_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05_split -s 1 -c 1 -o gpurun_out/r2c16_conv_block8 python profiles/microbench_split.py --levels 0 --shapes 96x96 --only fwd > gpurun_out/r2c16_ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tcgen05 -s 1 -c 1 -o gpurun_out/r2c16_wgrad_block8 python profiles/microbench_split.py --levels 0 --shapes 96x96 --only wgrad > gpurun_out/r2c16_ncu_wgrad.log 2>&1
ls -la gpurun_out | tail -12
