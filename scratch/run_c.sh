timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_nets.py -m gpu -x -q -k "conv_c3 or res_conv" 2>&1 | tail -3
DNB_THIN_TC_DGRAD=1 timeout 300 python bench.py --no-cpu --workload res_conv > gpurun_out/bench_res_conv_bf16.json 2> gpurun_out/bench_res_conv_bf16.err; echo "rc $?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_res_conv_bf16.json'))
print("TC dgrad", round(d['value']), d['unit'], d['ms_per_step'], d['final_loss'])
for t in d['top_kernels'][:24]:
    if t['kernel'].startswith('1:Conv2D'): print(f"   {t['ms_per_step']*1e3:8.1f} us {t['share']*100:5.1f}% {t['frac']} {t['kernel']}")
PY
