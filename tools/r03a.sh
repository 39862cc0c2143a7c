O=gpurun_out
for h in 0 1 2; do
echo "== L2HINT=$h" | tee -a $O/r03a.log
BGNN_GAT_L2HINT=$h python tools/bench_gat.py 20 20 64,128 2>&1 | tail -2 | tee -a $O/r03a.log
done
