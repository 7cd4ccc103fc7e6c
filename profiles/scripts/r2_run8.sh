#!/bin/bash
# Round 2, GPU call 8: GPU suite with the pinned-ReLU-decision gradient tests (reports: grad_pinned_*.json, c1_grad_report.json, loss_curve_100.json)
set -x
mkdir -p gpurun_out
export PCB_REPORT_DIR=$PWD/gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) 2>&1 | tail -40 > gpurun_out/r2c8_pytest.txt
