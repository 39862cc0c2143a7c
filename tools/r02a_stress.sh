set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
python tools/stress_smoke.py 2000 > gpurun_out/r02a_stress_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_stress_smoke.log
for i in $(seq 1 12); do BGNN_SMOKE_VERBOSE=1 python __graft_entry__.py smoke > gpurun_out/r02a_smoke_$i.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_smoke_$i.log; done
python tools/stress_knn.py 60 > gpurun_out/r02a_stress_knn.log 2>&1
tail -3 gpurun_out/r02a_pytest.log gpurun_out/r02a_stress_smoke.log gpurun_out/r02a_stress_knn.log
grep -L "rc=0" gpurun_out/r02a_smoke_*.log
