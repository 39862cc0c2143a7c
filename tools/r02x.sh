O=gpurun_out
export BGNN_F16_EW=2
timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -x -q 2>&1 | tail -2 | tee -a $O/r02x.log
for pr in 0 1; do
echo "== PAIR=$pr" | tee -a $O/r02x.log
BGNN_F16_PAIR=$pr BGNN_F16_DBG=8 python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "^cta" | grep -E "slot loads|3072 tiles" | sort | uniq | grep -E "warp 2:|warp 9:|issuer" | head -4 | tee -a $O/r02x.log
BGNN_F16_PAIR=$pr python tools/profile_knn.py f16 262144 786432 128 20 5 2>&1 | tail -1 | tee -a $O/r02x.log
BGNN_F16_PAIR=$pr ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed -k regex:knn_cosine_f16 --clock-control none -s 3 -c 1 \
    python tools/profile_knn.py f16 262144 786432 128 20 1 2>&1 | grep -E "gpu__time|tensor_cycles" | tee -a $O/r02x.log
done
BGNN_F16_PAIR=1 timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -x -q 2>&1 | tail -2 | tee -a $O/r02x.log
