O=gpurun_out
T=${1:-r02x2}
timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -x -q 2>&1 | tail -2 | tee -a $O/$T.log
for pr in 1; do
echo "== PAIR=$pr" | tee -a $O/$T.log
BGNN_F16_PAIR=$pr BGNN_F16_DBG=8 python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "^cta" | grep -E "slot loads|3072 tiles" | sort | uniq | grep -E "warp 2:|warp 9:|issuer" | head -3 | tee -a $O/$T.log
BGNN_F16_PAIR=$pr python tools/profile_knn.py f16 262144 786432 128 20 5 2>&1 | tail -1 | tee -a $O/$T.log
BGNN_F16_PAIR=$pr ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed -k regex:knn_cosine_f16 --clock-control none -s 3 -c 1 \
    python tools/profile_knn.py f16 262144 786432 128 20 1 2>&1 | grep -E "gpu__time|tensor_cycles" | tee -a $O/$T.log
done
