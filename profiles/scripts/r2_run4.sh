#!/bin/bash
# Round 2, GPU call 4: dual-format activation planes (fp16 forward / bf16 wgrad), PDL, rewritten column statistics + stem kernels,
# fused tcgen05 PointInfoNCE, fused inference executor.
set -x
mkdir -p gpurun_out
export PCB_REPORT_DIR=$PWD/gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c4_pytest.txt
PCB_PDL=1 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_c1.py tests/test_gpu_trainer.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2c4_pytest_pdl.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile-json gpurun_out/r2c4_profile.json > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
PCB_PDL=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c4_bench_pdl.json 2>> gpurun_out/r2c4_bench.err
PCB_FWD_FP16=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c4_bench_bf16.json 2>> gpurun_out/r2c4_bench.err
PCB_NCE_SIMT=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c4_bench_ncesimt.json 2>> gpurun_out/r2c4_bench.err
PCB_PDL=1 timeout 300 python bench.py --workload c0 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c4_bench_c0_pdl.json 2>> gpurun_out/r2c4_bench.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/r2c4_bench_c4.json 2>> gpurun_out/r2c4_bench.err
timeout 300 python profiles/grad_precision_ab.py 0.15 > gpurun_out/r2c4_grad_ab.json 2> gpurun_out/r2c4_grad_ab.log
ls -la gpurun_out | tail -12
