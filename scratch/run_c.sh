DNB_SV_PROF=1 python scratch/tw_bench.py 0 7 2>&1 | grep -E "tw prof|debug" | tail -5
