python scratch/prof_step.py mnist_cnn_wide bf16 > gpurun_out/prof_step_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_step_bf16_r01 -f python scratch/prof_step.py mnist_cnn_wide bf16 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
ncu --set full --clock-control none --profile-from-start off -k regex:maxpool_bwd_ring -o gpurun_out/prof_ps_r01 -f python scratch/prof_step.py mnist_cnn_wide bf16 per_sample > gpurun_out/ncu_ps.log 2>&1
tail -2 gpurun_out/ncu_ps.log
