set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_r01_default.json 2> gpurun_out/bench_r01_default.err; echo bench rc $?
timeout 600 python bench.py --impl reference > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_r01_reference.err; echo ref rc $?
cat gpurun_out/bench_r01_reference.json | cut -c1-600
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
python __graft_entry__.py smoke 2>&1 | tail -3
