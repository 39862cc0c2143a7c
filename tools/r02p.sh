( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02p_bench8.json 2> gpurun_out/r02p_bench8.err ) 2>&1 | tail -3
grep -v "Warning\|Consider\|loss0\|run_backward" gpurun_out/r02p_bench8.err | tail -c 1200
