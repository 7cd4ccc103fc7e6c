#!/bin/bash
# A/B sweep for the start of round 2 (one gpurun call, ~4 min): each line = one `bench.py` run of the C1 workload with one
# switch changed.  Usage on the GPU box:  bash profiles/r2_sweep.sh > gpurun_out/r2_sweep.txt
run() {
  local tag="$1"; shift
  local out
  out=$(env "$@" timeout 120 python bench.py --no-cpu-baseline --steps 8 --warmup 3 2>/dev/null | tail -1)
  python - "$tag" "$out" <<'PY'
import json, sys
tag, line = sys.argv[1], sys.argv[2]
try:
    l = json.loads(line); r = l["roofline"]; o = r["other"].get("wgrad_tcgen05_kernel", {})
    print(f'{tag:34s} {l["ms_per_step"]:7.2f} ms  {l["value"]:7.1f} pairs/s  e2e {l["e2e"]["value"]:7.1f}  conv {r["frac"]:.3f} of HBM '
          f'({r["avg_launch_ms"] * r["launches_per_step"]:.2f} ms)  wgrad {o.get("ms", 0):.2f} ms  loss {l["config"]["final_loss"]:.5f}')
except Exception as e:
    print(f"{tag:34s} FAILED ({e})")
PY
}
run "default" PCB_PAIR=1
run "two forward calls" PCB_PAIR=0
run "wgrad on a side stream" PCB_WGRAD_STREAM=1
run "conv split 0.25 wave" PCB_CONV_SPLIT_WAVES=0.25
run "conv split 1.0 wave" PCB_CONV_SPLIT_WAVES=1.0
run "wgrad split 0.5 wave" PCB_WGRAD_SPLIT_WAVES=0.5
run "wgrad split 2.0 waves" PCB_WGRAD_SPLIT_WAVES=2.0
run "tc5 cfg 1 (6 stages, 1 CTA/SM)" PCB_TC5_CFG=1
run "tc5 cfg 3 (2 stages, 3 CTAs/SM)" PCB_TC5_CFG=3
