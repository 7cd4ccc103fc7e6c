#!/bin/bash
# Round 2, GPU call 1: new C1-shape parity tests + full GPU suite, gradient-precision A/B, baseline bench + A/B sweep,
# ncu launch list of the stacked schedule (with DRAM bytes), ncu --set full of the weight-gradient kernel on two shapes.
set -x
mkdir -p gpurun_out
export PCB_REPORT_DIR=$PWD/gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_gpu.txt
python -c "import os; print('cores', len(os.sched_getaffinity(0)))" >> gpurun_out/r2c1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2c1_pytest.txt
timeout 300 python profiles/grad_precision_ab.py 0.15 > gpurun_out/r2c1_grad_ab.json 2> gpurun_out/r2c1_grad_ab.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
timeout 600 bash profiles/r2_sweep.sh > gpurun_out/r2c1_sweep.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2700 -c 1100 --csv \
  --log-file gpurun_out/r2c1_launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2c1_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tcgen05 -s 1 -c 1 -o gpurun_out/r2c1_wgrad_block8 \
  python profiles/microbench_split.py --levels 0 --shapes 96x96 --only wgrad > gpurun_out/r2c1_ncu_wg0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tcgen05 -s 1 -c 1 -o gpurun_out/r2c1_wgrad_s16 \
  python profiles/microbench_split.py --levels 4 --shapes 256x256 --only wgrad > gpurun_out/r2c1_ncu_wg4.log 2>&1
timeout 300 python profiles/microbench_split.py --levels 0,1,2,3,4 --shapes 96x96,128x128,256x256 > gpurun_out/r2c1_microbench.txt 2>&1
ls -la gpurun_out | tail -30
