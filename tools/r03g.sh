O=gpurun_out
export BGNN_F16_TSPLIT=1
timeout 900 python -m pytest tests/test_gpu_knn.py -m gpu -x -q 2>&1 | tail -2 | tee $O/r03g.log
for ts in 1 0; do
echo "== TSPLIT=$ts" | tee -a $O/r03g.log
BGNN_F16_TSPLIT=$ts BGNN_F16_DBG=8 python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "^cta" | grep -E "slot loads|3072 tiles" | sort | uniq | grep -E "warp 2:|warp 9:|issuer" | head -3 | cut -c1-260 | tee -a $O/r03g.log
BGNN_F16_TSPLIT=$ts python tools/profile_knn.py f16 262144 786432 128 20 5 2>&1 | tail -1 | tee -a $O/r03g.log
BGNN_F16_TSPLIT=$ts ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed -k regex:knn_cosine_f16 --clock-control none -s 3 -c 1 \
    python tools/profile_knn.py f16 262144 786432 128 20 1 2>&1 | grep -E "gpu__time|tensor_cycles" | tee -a $O/r03g.log
done
BGNN_F16_TSPLIT=1 timeout 300 python tools/stress_knn.py 4 2>&1 | tail -1 | tee -a $O/r03g.log
