O=gpurun_out
T=${1:-r02x4}
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_assemble.py -m gpu -x -q 2>&1 | tail -2 | tee -a $O/$T.log
python tools/profile_knn.py f16 262144 786432 128 20 5 2>&1 | tail -1 | tee -a $O/$T.log
ncu --metrics gpu__time_duration.sum -k regex:knn_ --clock-control none -s 8 -c 6 python tools/profile_knn.py f16 262144 786432 128 20 1 2>&1 | grep -E "^  [a-z_ ]*knn_[a-z_0-9]+|gpu__time" | sed 's/(.*//' | tee -a $O/$T.log
timeout 300 python tools/stress_knn.py 4 2>&1 | tail -1 | tee -a $O/$T.log
