timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_nets.py -m gpu -x -q -k "conv or cnn" 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --no-alt > gpurun_out/bench_c.json 2>gpurun_out/bench_c.err; echo rc $?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['final_loss'])
for t in d['top_kernels'][:5]: print(f"{t['ms_per_step']*1e3:8.1f} us {t['frac']} {t['kernel']}")
PY
