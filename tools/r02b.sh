python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -30 gpurun_out/r02b_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02b_bench.err
