O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -m gpu -x -q 2>&1 | tail -2 | tee -a $O/r02y.log
for pr in 1 0; do
echo "== config-5 shape sample (65536 x 12582912, d=256, k=32) PAIR=$pr" | tee -a $O/r02y.log
BGNN_F16_PAIR=$pr timeout 600 python tools/profile_knn.py f16 65536 12582912 256 32 2 2>&1 | tail -1 | tee -a $O/r02y.log
echo "== config-4 PAIR=$pr" | tee -a $O/r02y.log
BGNN_F16_PAIR=$pr timeout 600 python tools/profile_knn.py f16 262144 786432 128 20 5 2>&1 | tail -1 | tee -a $O/r02y.log
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-sync16m > $O/r02y_bench.json 2>$O/r02y_bench.err; python - <<'PY' | tee -a gpurun_out/r02y.log
import json
d=json.loads(open('gpurun_out/r02y_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], 'knn', d['knn_build']['ms'], d['knn_build']['roofline']['frac'], d['knn_build']['e2e'], d['knn_build']['exact_fallback_rows'])
PY
