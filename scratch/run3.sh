python scratch/prof_step.py > gpurun_out/prof_step_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_times.csv python scratch/prof_step.py > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
