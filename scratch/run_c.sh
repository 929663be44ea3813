timeout 600 python bench.py > gpurun_out/bench_r01_default.json 2> gpurun_out/bench_r01_default.err; echo bench rc $?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r01_default.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], d['other_maxpool_route'], d['roofline']['kernel'], d['roofline']['frac'])
PY
tail -3 gpurun_out/bench_r01_default.err
