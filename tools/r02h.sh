python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
tail -12 gpurun_out/r02h_pytest.log | cut -c1-600
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-sync16m > gpurun_out/r02h_bench2.json 2> gpurun_out/r02h_bench2.err ) 2>&1 | tail -3
grep -v Warning gpurun_out/r02h_bench2.err | tail -c 1500
