#!/bin/bash
# Round 2, GPU call 6: full GPU suite on the cleaned-up library, default bench (outlier check), launch list with DRAM bytes, ncu --set full
# of the dominant kernel in the benchmarked schedule.
set -x
mkdir -p gpurun_out
export PCB_REPORT_DIR=$PWD/gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2c6_pytest.txt
timeout 300 python bench.py --steps 50 --warmup 5 --profile-json gpurun_out/r2c6_profile.json > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c6_bench_20.json 2>> gpurun_out/r2c6_bench.err
timeout 300 python bench.py --loss hardest --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c6_bench_hardest.json 2>> gpurun_out/r2c6_bench.err
timeout 300 python bench.py --workload c0 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c6_bench_c0.json 2>> gpurun_out/r2c6_bench.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/r2c6_bench_c4.json 2>> gpurun_out/r2c6_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c6_bench_reference.json 2>> gpurun_out/r2c6_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c6_smoke.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2500 -c 1700 --csv \
  --log-file gpurun_out/r2c6_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c6_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05_split -s 1 -c 1 -o gpurun_out/r2c6_conv_block8 \
  python profiles/microbench_split.py --levels 0 --shapes 96x96 --only fwd > gpurun_out/r2c6_ncu_conv.log 2>&1
ls -la gpurun_out | tail -12
