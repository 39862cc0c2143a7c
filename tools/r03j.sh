O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mp.py -m gpu -x -q 2>&1 | tail -2 | tee $O/r03j.log
python tools/profile_mp.py 20 > $O/r03j_profile_mp.log 2>&1; python tools/prof_table.py $O/r03j_profile_mp.log 12 | tee -a $O/r03j.log
