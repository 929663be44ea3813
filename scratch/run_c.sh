timeout 300 python bench.py --no-cpu --math bf16 > gpurun_out/bench_c.json 2>gpurun_out/bench_c.err; echo rc $?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'])
tot=0
for t in d['top_kernels']:
    tot+=t['ms_per_step']; print(f"{t['ms_per_step']*1e3:8.1f} us {t['share']*100:5.1f}% {t['kernel']}")
print('sum top', tot)
PY
