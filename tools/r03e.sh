N=${1:-2}
python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r03e_pytest$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03e_pytest$N.log
tail -4 gpurun_out/r03e_pytest$N.log | cut -c1-600
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r03e_bench$N.json 2> gpurun_out/r03e_bench$N.err ) 2>&1 | tail -3
grep -v "Warning\|Consider\|loss0\|run_backward" gpurun_out/r03e_bench$N.err | tail -c 800
