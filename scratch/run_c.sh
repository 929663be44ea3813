for w in mnist_cnn_lit mnist_mlp res_conv rnn_sentiment; do
timeout 200 python bench.py --no-cpu --no-alt --steps 10 --workload $w > gpurun_out/bench_${w}_bf16.json 2> gpurun_out/bench_${w}_bf16.err; echo "$w rc $?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${w}_bf16.json'))
print('  ', round(d['value']), d['unit'], round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'loss', d['final_loss'])
PY
done
