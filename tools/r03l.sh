O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mp.py tests/test_gpu_assemble.py -m gpu -x -q 2>&1 | tail -3 | tee $O/r03l.log
python tools/bench_gat.py 20 20 64,128,2,32 2>&1 | tail -4 | tee -a $O/r03l.log
python tools/profile_mp.py 20 > $O/r03l_profile_mp.log 2>&1; python tools/prof_table.py $O/r03l_profile_mp.log 10 | tee -a $O/r03l.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-80 | tee -a $O/r03l.log
