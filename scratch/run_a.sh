set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo bench rc $?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_a_ref.json 2> gpurun_out/bench_a_ref.err; echo ref rc $?
python scratch/prof_step.py > gpurun_out/prof_step_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_step_a -f python scratch/prof_step.py > gpurun_out/ncu_step.log 2>&1
tail -3 gpurun_out/ncu_step.log
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_a.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
