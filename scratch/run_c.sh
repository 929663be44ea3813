timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_nets.py -m gpu -x -q -k "maxpool or per_sample or route" 2>&1 | tail -4
timeout 300 python bench.py --no-cpu --route per_sample > gpurun_out/bench_ps_bf16.json 2>gpurun_out/bench_ps_bf16.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ps_bf16.json'))
print('per_sample', round(d['value']), d['ms_per_step'])
for t in d['top_kernels'][:10]: print(f"   {t['ms_per_step']*1e3:8.1f} us {t['share']*100:5.1f}% {t['frac']} {t['kernel']}")
PY
