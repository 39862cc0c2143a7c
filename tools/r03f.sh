O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mp.py -m gpu -x -q -k "gcn or sage or graph_prepare or fast_graph" 2>&1 | tail -3 | tee $O/r03f.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a $O/r03f.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-sync16m > $O/r03f_bench.json 2>$O/r03f_bench.err; python - <<'PY' | tee -a gpurun_out/r03f.log
import json
d=json.loads(open('gpurun_out/r03f_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['frac'], 'knn', d['knn_build']['ms'], d['knn_build']['roofline']['frac'], 'knn e2e', d['knn_build']['e2e']['ms'], d['knn_build']['exact_fallback_rows'])
PY
