# kNN sweep, CTA pairs (cta_group::2) with the round-2 epilogue: modes + wait cycles
O=gpurun_out
export BGNN_F16_PAIR=1
export BGNN_F16_EW=2
for m in 0 2 6; do
  echo "== PAIR mode $m (nq=37888)" | tee -a $O/r02t.log
  BGNN_F16_DBG=$m ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__cycles_active.avg.per_second,lts__throughput.avg.pct_of_peak_sustained_elapsed \
    -k regex:knn_cosine_f16 --clock-control none -s 1 -c 1 python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "gpu__time|tensor_cycles|per_second|lts__" | tee -a $O/r02t.log
done
for m in 8 10 14; do
echo "== PAIR waits dbg=$m" | tee -a $O/r02t.log
BGNN_F16_DBG=$m python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "^cta" | grep -E "slot loads|3072 tiles" | sort | uniq | head -30 | tee -a $O/r02t.log
done
export BGNN_F16_PAIR=0
for m in 0 2 6; do
  echo "== SINGLE mode $m (nq=37888) with lts" | tee -a $O/r02t.log
  BGNN_F16_DBG=$m ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum \
    -k regex:knn_cosine_f16 --clock-control none -s 1 -c 1 python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "gpu__time|tensor_cycles|lts__" | tee -a $O/r02t.log
done
echo "== SINGLE waits dbg=10 (no epilogue)" | tee -a $O/r02t.log
BGNN_F16_DBG=10 python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "^cta" | grep -E "slot loads|3072 tiles" | sort | uniq | head -8 | tee -a $O/r02t.log
