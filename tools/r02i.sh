( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-sync16m > gpurun_out/r02i_bench2.json 2> gpurun_out/r02i_bench2.err ) 2>&1 | tail -3
grep -v Warning gpurun_out/r02i_bench2.err | tail -c 800
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/dist_spmm_check.py 22 256 > gpurun_out/r02i_spmm.log 2>&1; grep "group:" gpurun_out/r02i_spmm.log || tail -20 gpurun_out/r02i_spmm.log
