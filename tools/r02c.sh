python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -40 gpurun_out/r02c_pytest.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for v in 0 1 2 3 4; do
  BGNN_GAT_VARIANT=$v ncu --metrics $M -k regex:gatv2_ --clock-control none --csv --log-file gpurun_out/r02c_A$v.csv python tools/bench_gat.py 20 1 64 > gpurun_out/r02c_A$v.log 2>&1
done
for v in 1 2; do
  BGNN_GAT_BVARIANT=$v ncu --metrics $M -k regex:gatv2_bwd_src --clock-control none --csv --log-file gpurun_out/r02c_B$v.csv python tools/bench_gat.py 20 1 64 > gpurun_out/r02c_B$v.log 2>&1
done
ncu --set full --import-source on -k regex:gatv2_bwd --clock-control none -c 6 -o gpurun_out/r02c_gat_full python tools/bench_gat.py 20 1 64 > gpurun_out/r02c_full.log 2>&1
python tools/bench_gat.py 20 10 64,128 > gpurun_out/r02c_bench_gat.log 2>&1; cat gpurun_out/r02c_bench_gat.log
