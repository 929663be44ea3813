timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_layers.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu > gpurun_out/bench_c.json 2>gpurun_out/bench_c.err; echo rc $?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['final_loss'], d['gpu_launches'], d['other_maxpool_route']['ms_per_step'])
for t in d['top_kernels'][:24]:
    if 'MaxPool2D:fwd' in t['kernel'] or 'Relu:fwd' in t['kernel']: print(f"{t['ms_per_step']*1e3:8.1f} us {t['frac']} {t['kernel']}")
PY
