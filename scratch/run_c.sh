timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_layers.py tests/test_abi.py -m gpu -x -q -k "nets or maxpool or abi or trajectory or curve or fit or mnist" 2>&1 | tail -4
timeout 300 python bench.py --no-cpu > gpurun_out/bench_c.json 2>gpurun_out/bench_c.err; echo rc $?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['final_loss'], d['gpu_launches'], d['other_maxpool_route'])
PY
tail -3 gpurun_out/bench_c.err
