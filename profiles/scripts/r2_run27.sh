#!/bin/bash
# Round 2, GPU call 27: TMA tile::gather4 probe (stand-alone binary, profiles/probes/gather4_probe.cu): landing layout, zero fill, sustained rate.
mkdir -p gpurun_out
P=profiles/probes/gather4_probe
{
for cfg in "layout 1 0" "layout 4 0" "layout 1 64" "rate 1 0 128" "rate 1 64 128" "rate 1 0 32"; do
  timeout 20 $P $cfg 2>&1 | head -60
  echo "-- exit $?"
done
} > gpurun_out/r2c27_gather4.txt 2>&1
tail -5 gpurun_out/r2c27_gather4.txt
