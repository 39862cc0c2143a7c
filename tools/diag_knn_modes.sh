#!/bin/bash
# Bottleneck experiments for knn_cosine_f16_kernel: BGNN_F16_DBG bit0 = no selection, bit1 = no tcgen05.ld,
# bit2 = no MMA.  Kernel durations from ncu (results are garbage in modes != 0).  $2 = BGNN_F16_PAIR.
set -u
NQ=${1:-37888}
export BGNN_F16_PAIR=${2:-1}
for m in 0 1 2 4 6; do
  echo "== mode $m pair $BGNN_F16_PAIR"
  BGNN_F16_DBG=$m ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__cycles_active.avg.per_second \
    -k regex:knn_cosine_f16 --clock-control none -s 1 -c 1 python tools/profile_knn.py f16 $NQ 786432 128 20 1 2>&1 | grep -E "gpu__time|tensor_cycles|per_second"
done
