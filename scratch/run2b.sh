for d in 0 1 2 4 3 7; do DNB_WG_DEBUG=$d python scratch/sv_bench.py 0 2>&1 | tail -1 | sed "s/.*tc_conv2d_bwd_weight_sv/wg debug $d:/"; done
