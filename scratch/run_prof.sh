set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r01_default.json 2> gpurun_out/bench_r01_default.err; echo bench rc $?
python bench.py --math fp32 --no-cpu > gpurun_out/bench_r01_fp32.json 2> gpurun_out/bench_r01_fp32.err; echo bench rc $?
python scratch/prof_step.py > gpurun_out/prof_step_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_step_r01 -f python scratch/prof_step.py > gpurun_out/ncu_step.log 2>&1
tail -3 gpurun_out/ncu_step.log
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
