python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -5 gpurun_out/r02d_pytest.log
python tools/debug_fb.py > gpurun_out/r02d_debug_fb.log 2>&1; head -5 gpurun_out/r02d_debug_fb.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for v in 0 5 2; do
  BGNN_GAT_VARIANT=$v ncu --metrics $M -k regex:gatv2_bwd_dst --clock-control none --csv --log-file gpurun_out/r02d_A$v.csv python tools/bench_gat.py 20 1 64 > gpurun_out/r02d_A$v.log 2>&1
done
for v in 0 2; do
  BGNN_GAT_BVARIANT=$v ncu --metrics $M -k regex:gatv2_bwd_src --clock-control none --csv --log-file gpurun_out/r02d_B$v.csv python tools/bench_gat.py 20 1 64 > gpurun_out/r02d_B$v.log 2>&1
done
python tools/bench_gat.py 20 10 64,128 > gpurun_out/r02d_bench_gat.log 2>&1; cat gpurun_out/r02d_bench_gat.log
