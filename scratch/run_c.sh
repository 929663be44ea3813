timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu --no-alt > gpurun_out/bench_c.json 2>gpurun_out/bench_c.err; echo rc $?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['final_loss'], d['gpu_launches'])
PY
