python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
tail -6 gpurun_out/r02m_pytest.log | cut -c1-1200
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02m_bench4.json 2> gpurun_out/r02m_bench4.err ) 2>&1 | tail -3
grep -v "Warning\|Consider\|loss0\|run_backward" gpurun_out/r02m_bench4.err | tail -c 800
