timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -x -q -k "conv or Conv" 2>&1 | tail -6
timeout 120 python scratch/sv_bench.py 0 2>&1 | tail -2
timeout 300 python bench.py --no-cpu --steps 10 > gpurun_out/b2.json 2> gpurun_out/b2.err; tail -3 gpurun_out/b2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b2.json'))
print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])
print([(k['kernel'].split('/')[-1], round(k['ms_per_step'],3)) for k in d['top_kernels']])
PY
