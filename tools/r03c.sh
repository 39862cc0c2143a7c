O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r03c_pytest.log 2>&1; tail -4 $O/r03c_pytest.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r03c_bench.json 2> $O/r03c_bench.err ) 2> $O/r03c_bench.time; cat $O/r03c_bench.time; tail -2 $O/r03c_bench.err
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r03c_bench_reference.json 2> $O/r03c_bench_reference.err ) 2> $O/r03c_ref.time; cat $O/r03c_ref.time; tail -c 600 $O/r03c_bench_reference.json
