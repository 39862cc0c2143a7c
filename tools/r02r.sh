# kNN sweep: four epilogue warps per TMEM lane quarter (EW=4, default) against two (BGNN_F16_EW=2)
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -x -q 2>&1 | tail -4 > $O/r02r_pytest.log; cat $O/r02r_pytest.log
for ew in 4 2; do
  echo "== EW=$ew" | tee -a $O/r02r_knn.log
  BGNN_F16_EW=$ew timeout 300 python tools/profile_knn.py f16 262144 786432 128 20 5 2>&1 | tail -1 | tee -a $O/r02r_knn.log
done
BGNN_F16_EW=2 timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -x -q -k "bit_identical or uncertified or sync_1m" 2>&1 | tail -2 | tee -a $O/r02r_pytest.log
timeout 300 python tools/stress_knn.py 6 2>&1 | tail -2 | tee -a $O/r02r_knn.log
