for n in 8; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/b_dp$n.json 2> gpurun_out/b_dp$n.err; echo rc $?
python - <<PY
import json
d=json.load(open('gpurun_out/b_dp$n.json'))
print(d['n_gpus'], round(d['value']), d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'], d['gpu_launches'])
PY
done
