O=gpurun_out
for dr in 8 4 16 32; do
echo "== DRAIN=$dr" | tee -a $O/r03h.log
BGNN_F16_DRAIN=$dr ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed -k regex:knn_cosine_f16 --clock-control none -s 3 -c 1 \
    python tools/profile_knn.py f16 262144 786432 128 20 1 2>&1 | grep -E "gpu__time|tensor_cycles|fallback" | tee -a $O/r03h.log
done
