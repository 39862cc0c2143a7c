python -m pytest tests/test_gpu_mp.py -m gpu -q -k "heads or ktgnn" 2>&1 | tail -2
python tools/profile_knn.py f16 262144 786432 128 20 3 2>&1 | tail -2
BGNN_F16_BN=128 python tools/profile_knn.py f16 262144 786432 128 20 3 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-sync16m > gpurun_out/r02o_bench.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02o_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['step_cuda_graph'], d['e2e']['ms_per_step'])
print(d['roofline']['kernel_ms_per_step'])
PY
