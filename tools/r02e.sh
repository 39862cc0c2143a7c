python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -4 gpurun_out/r02e_pytest.log
( time python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02e_bench.err
nvidia-smi --query-gpu=memory.used --format=csv
