#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k batched_weight 2>&1 | tail -40 > gpurun_out/r2c18_pytest.txt
