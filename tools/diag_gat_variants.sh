#!/bin/bash
# Pass-A tuning variants (BGNN_GAT_VARIANT) on the c = 64 mapping: kernel durations from ncu.
for v in 0 1 2 3 4; do
  echo "== variant $v"
  BGNN_GAT_VARIANT=$v ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active -k regex:gatv2_bwd --clock-control none python tools/bench_gat.py 20 1 64 2>&1 | grep -E "gatv2_bwd|gpu__time|inst_executed|warps_active" | awk 'NR>12'
done
