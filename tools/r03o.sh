O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mp.py -m gpu -x -q 2>&1 | tail -3 | tee $O/r03o.log
python tools/bench_gat.py 20 20 64,128,32 2>&1 | tail -3 | tee -a $O/r03o.log
BGNN_GAT_NOSPLIT=1 python tools/bench_gat.py 20 20 64 2>&1 | tail -1 | tee -a $O/r03o.log
python tools/profile_mp.py 20 > $O/r03o_profile_mp.log 2>&1; python tools/prof_table.py $O/r03o_profile_mp.log 8 | tee -a $O/r03o.log
