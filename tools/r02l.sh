python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log
tail -6 gpurun_out/r02l_pytest.log | cut -c1-1200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/dist_spmm_check.py 22 256 > gpurun_out/r02l_spmm.log 2>&1; grep "group:" gpurun_out/r02l_spmm.log || tail -20 gpurun_out/r02l_spmm.log
