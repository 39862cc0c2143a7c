O=gpurun_out
( for i in 1 2; do timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done ) | tee $O/r03i_stress.log
timeout 900 python tools/stress_smoke.py 600 2>&1 | tail -3 | tee -a $O/r03i_stress.log
timeout 600 python tools/stress_knn.py 30 2>&1 | tail -2 | tee -a $O/r03i_stress.log
for i in 1 2 3 4 5 6; do python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-60; done | sort | uniq -c | tee -a $O/r03i_stress.log
