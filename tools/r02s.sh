# kNN sweep diagnostics for EW = 2 / 4: per-kernel times of a full build, bottleneck modes, wait cycles
O=gpurun_out
export BGNN_F16_PAIR=0
for ew in 2 4; do
  export BGNN_F16_EW=$ew
  echo "==== EW=$ew: kernels of one full build (262144 x 786432)" | tee -a $O/r02s.log
  ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed -k regex:knn_ --clock-control none -s 10 -c 5 \
    python tools/profile_knn.py f16 262144 786432 128 20 1 2>&1 | grep -E "knn_[a-z_0-9]+|gpu__time|tensor_cycles" | sed 's/(CUtensor.*//' | tee -a $O/r02s.log
  for m in 0 1 2 4; do
    echo "== EW=$ew mode $m (nq=37888)" | tee -a $O/r02s.log
    BGNN_F16_DBG=$m ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__cycles_active.avg.per_second \
      -k regex:knn_cosine_f16 --clock-control none -s 1 -c 1 python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "gpu__time|tensor_cycles|per_second" | tee -a $O/r02s.log
  done
  echo "== EW=$ew waits" | tee -a $O/r02s.log
  BGNN_F16_DBG=8 python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "^cta" | grep -E "3072 slot|3072 tiles" | sort | uniq | head -24 | tee -a $O/r02s.log
done
