O=gpurun_out
for h in "" 1; do
echo "== SEQ_RECORDS='$h'" | tee -a $O/r03k.log
BGNN_HACK_SEQ_RECORDS=$h python tools/bench_gat.py 20 20 64,2 2>&1 | tail -2 | tee -a $O/r03k.log
done
