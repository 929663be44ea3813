python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scratch/sv_bench.py 0 1 2 4 3 5 6 7 2>&1 | tail -12
python bench.py --math tf32 > gpurun_out/bench_tf32_a.json 2> gpurun_out/bench_tf32_a.err; tail -c 600 gpurun_out/bench_tf32_a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_tf32_a.json'))
print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['cpu_baseline'])
print([(k['kernel'], round(k['ms_per_step'],3)) for k in d['top_kernels']])
PY
