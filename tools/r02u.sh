# kNN sweep, CTA pairs after the remote-arrive fix (no cluster-scope release): modes, waits, full build, parity
O=gpurun_out
export BGNN_F16_EW=2
export BGNN_F16_PAIR=1
for m in 0 2; do
  echo "== PAIR mode $m (nq=37888)" | tee -a $O/r02u.log
  BGNN_F16_DBG=$m ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__cycles_active.avg.per_second,lts__throughput.avg.pct_of_peak_sustained_elapsed \
    -k regex:knn_cosine_f16 --clock-control none -s 1 -c 1 python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "gpu__time|tensor_cycles|per_second|lts__" | tee -a $O/r02u.log
done
for m in 8 10; do
echo "== PAIR waits dbg=$m" | tee -a $O/r02u.log
BGNN_F16_DBG=$m python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "^cta" | grep -E "slot loads|3072 tiles" | sort | uniq | grep -E "warp 2:|warp 5:|issuer|producer" | head -12 | tee -a $O/r02u.log
done
for pr in 1 0; do
  echo "== full build PAIR=$pr" | tee -a $O/r02u.log
  BGNN_F16_PAIR=$pr python tools/profile_knn.py f16 262144 786432 128 20 5 2>&1 | tail -1 | tee -a $O/r02u.log
  BGNN_F16_PAIR=$pr ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed -k regex:knn_cosine_f16 --clock-control none -s 3 -c 1 \
    python tools/profile_knn.py f16 262144 786432 128 20 1 2>&1 | grep -E "gpu__time|tensor_cycles" | tee -a $O/r02u.log
done
BGNN_F16_PAIR=1 timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -x -q 2>&1 | tail -2 | tee -a $O/r02u.log
BGNN_F16_PAIR=1 timeout 300 python tools/stress_knn.py 4 2>&1 | tail -1 | tee -a $O/r02u.log
