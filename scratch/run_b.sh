set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for m in bf16; do
timeout 300 python bench.py --no-cpu --math $m > gpurun_out/bench_b_$m.json 2> gpurun_out/bench_b_$m.err; echo bench rc $?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_b_$m.json'))
print('$m', d['value'], d['ms_per_step'], d['e2e']['value'], d['final_loss'])
for t in d['top_kernels']: print(f"{t['ms_per_step']*1e3:8.1f} us {t['share']*100:5.1f}% {t['bound']} {t['frac']} {t['kernel']}")
PY
done
