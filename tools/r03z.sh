O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $O/r03z.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-60 | tee -a $O/r03z.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r03z_bench.json 2>$O/r03z_bench.err; python - <<'PY' | tee -a gpurun_out/r03z.log
import json
d=json.loads(open('gpurun_out/r03z_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], 'knn', d['knn_build']['ms'], d['knn_build']['roofline']['frac'], d['knn_build']['roofline']['traffic'], 'knn e2e', d['knn_build']['e2e']['ms'], 'sync16m build', d['sync16m']['knn_build']['ms'], d['clocks'])
PY
