python -m pytest tests/test_gpu_mp.py -m gpu -q -k "gat or ktgnn or adapted_conv" 2>&1 | tail -3
python tools/bench_gat.py 20 10 64,128,32 2>&1 | tail -3
BGNN_GAT_HUBS=separate python tools/bench_gat.py 20 10 64 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-sync16m > gpurun_out/r02q_bench.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['step_cuda_graph'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])
PY
