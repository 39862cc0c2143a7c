O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mp.py -m gpu -x -q 2>&1 | tail -3 | tee -a $O/r03b.log
python tools/profile_e2e.py 2>&1 | grep -E "prepare_graph:|RadixSort|prep_|index_elementwise|Self CUDA time" | cut -c1-160 | tee -a $O/r03b.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-sync16m > $O/r03b_bench.json 2>$O/r03b_bench.err; python - <<'PY' | tee -a gpurun_out/r03b.log
import json
d=json.loads(open('gpurun_out/r03b_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['frac'], 'knn', d['knn_build']['ms'], d['knn_build']['roofline']['frac'], d['knn_build']['e2e']['ms'], d['knn_build']['exact_fallback_rows'])
PY
tail -3 $O/r03b_bench.err
