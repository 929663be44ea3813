python scratch/sv_bench.py 0 > gpurun_out/svb_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tc_conv_sv_kernel|wgrad_sv" -s 5 -c 5 -o gpurun_out/prof_sv3_r01 -f python scratch/sv_bench.py 0 > gpurun_out/ncu_sv3.log 2>&1
tail -2 gpurun_out/ncu_sv3.log
